// Native sm_100a FFT back end of the IMEX plan (power-of-two extents).  Placeholder until
// the pass kernels land: reports "unsupported" so that EVX_FFT_AUTO selects cuFFT.
#include "evx_internal.h"
#include "spectral_plan.h"

namespace evx {

bool native_fft_supported(int, int, int) { return false; }
int native_plan_init(evx_imex_plan*) { return EVX_ERR_UNSUPPORTED; }
void native_plan_free(evx_imex_plan* p) {
  if (p && p->twiddles) { cudaFree(p->twiddles); p->twiddles = nullptr; }
}
int native_apply(evx_imex_plan*, const float*, const float*, float*, void*, const double*, double,
                 double, int, cudaStream_t) { return EVX_ERR_UNSUPPORTED; }
int native_ch_step(evx_imex_plan*, const float*, const float*, float*, void*, const double*,
                   double, double, double, double, cudaStream_t) { return EVX_ERR_UNSUPPORTED; }

}  // namespace evx
