// stand-in for <pcl/pcl_base.h>: see oracle/shim/README.md
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
