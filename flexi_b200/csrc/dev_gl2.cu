// dev_gl2.cu -- development aid: instantiates only the headline k_volume_gl2 (N=7, Pirozzoli, Navier-Stokes, nVarLift 5 with
// 4 lifted variables, BR1) so that its registers / SASS show in seconds instead of rebuilding every degree.  Not part of the library.
#include "volume_gl2.cuh"
template __global__ void fx::k_volume_gl2<8, FLEXI_SPLITDG_PI, true, 5, 4, 1>(const __grid_constant__ fx::VolArgs2<8>);
