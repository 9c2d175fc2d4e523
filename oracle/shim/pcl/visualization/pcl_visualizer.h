// Do-nothing stand-ins for <pcl/visualization/pcl_visualizer.h> and the VTK classes that
// Graph::voxelViewer (graph.hpp:67-129) names: the method has to compile, it is never called.
// See oracle/shim/README.md.
#ifndef IFTWT_SHIM_PCL_VISUALIZER_H
#define IFTWT_SHIM_PCL_VISUALIZER_H
#include <pcl/point_cloud.h>
#include <string>
template <class T>
class vtkSmartPointer {
 public:
  static vtkSmartPointer New() { return vtkSmartPointer(); }
  T* operator->() const { static T t; return &t; }
  operator T*() const { static T t; return &t; }
};
struct vtkShimAny {
  template <class... A> void SetBounds(A&&...) {}
  template <class... A> void Update(A&&...) {}
  template <class... A> void AddInputData(A&&...) {}
  template <class... A> void SetInputConnection(A&&...) {}
  template <class... A> void SetMapper(A&&...) {}
  template <class... A> void SetColor(A&&...) {}
  template <class... A> void SetAmbient(A&&...) {}
  template <class... A> void SetLineWidth(A&&...) {}
  template <class... A> void SetOpacity(A&&...) {}
  template <class... A> void AddActor(A&&...) {}
  void EdgeVisibilityOn() {}
  void SetRepresentationToWireframe() {}
  vtkShimAny* GetOutput() { return this; }
  vtkShimAny* GetOutputPort() { return this; }
  vtkShimAny* GetProperty() { return this; }
  vtkShimAny* GetRenderers() { return this; }
  vtkShimAny* GetFirstRenderer() { return this; }
};
struct vtkAppendPolyData : vtkShimAny {};
struct vtkCubeSource : vtkShimAny {};
struct vtkCleanPolyData : vtkShimAny {};
struct vtkPolyDataMapper : vtkShimAny {};
struct vtkActor : vtkShimAny {};
struct vtkRenderWindow : vtkShimAny {};
namespace pcl {
namespace visualization {
enum RenderingProperties { PCL_VISUALIZER_POINT_SIZE };
class PCLVisualizer {
 public:
  explicit PCLVisualizer(const std::string& = "") {}
  vtkSmartPointer<vtkRenderWindow> getRenderWindow() { return vtkSmartPointer<vtkRenderWindow>(); }
  template <class PointT, class CloudPtr>
  bool addPointCloud(const CloudPtr&, const std::string& = "cloud") { return true; }
  void setBackgroundColor(double, double, double) {}
  void setPointCloudRenderingProperties(int, double, const std::string& = "cloud") {}
  void setCameraPosition(double, double, double, double, double, double) {}
};
}  // namespace visualization
}  // namespace pcl
#endif
