// Second translation unit of a3d_wave.cu: the SimpleRayCaster instantiations of k_cast_wave (compiled
// in parallel with the IterativeRayCaster ones; see the note at the top of a3d_wave.cu).
#define A3D_WAVE_PART 2
#include "a3d_wave.cu"
