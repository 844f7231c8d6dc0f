// Second translation unit of a3d_split.cu: the DIRECT instantiations of the mark / leaf kernels (second
// pass of a stored-list call, launch_cast_split_direct), compiled in parallel with the first.
#define A3D_SPLIT_PART 2
#include "a3d_split.cu"
