// stand-in for <pcl/io/pcd_io.h> (file I/O is not on the path the driver runs): see oracle/shim/README.md
#include <pcl/point_cloud.h>
