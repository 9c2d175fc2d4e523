// stand-in, see pcl/visualization/pcl_visualizer.h in this directory
#include <pcl/visualization/pcl_visualizer.h>
