// Shim: stands in for <voxblox_ros/esdf_server.h> (oracle/_ref, test infrastructure only).
#include "../voxblox_shim.h"
