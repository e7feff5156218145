// Plan object behind evx_imex_plan_* and declarations shared by spectral.cu / fft_native.cu.
#pragma once
#include <cufft.h>
#include <cuda_runtime.h>
#include "evx_hd.h"
#include "spectral_math.h"
#include "../../include/evoxels_b200.h"

struct evx_imex_plan {
  int nx = 0, ny = 0, nz = 0, is_f64 = 0, backend = 0;
  int n[3] = {1, 1, 1};        // extents with size-1 axes squeezed to the front
  int axis_of[3] = {-1, -1, -1};  // which original axis each squeezed axis is
  int rank = 0;
  size_t real_elems = 0, spec_elems = 0;
  size_t real_bytes = 0, spec_bytes = 0, work_bytes = 0;
  bool have_cufft = false;
  cufftHandle fwd = 0, inv = 0;
  // native back end
  void* twiddles = nullptr;    // device table(s), owned
  int spec_pitch = 0;          // complex elements per (x,y) row of the native spectrum
  // TMA tensor maps of the native spectrum (y-pass and x-pass boxes), encoded on first use
  // for the workspace address `tmap_spec` (fft_line.cu)
  alignas(64) unsigned char tmap_y[128];
  alignas(64) unsigned char tmap_x[128];
  void* tmap_spec = nullptr;
  int tmap_kz = 0;
  // tensor map of the chained z/y passes ([8 x 256 x 1] boxes along y), same keying
  alignas(64) unsigned char tmap_chain[128];
  void* tmap_chain_spec = nullptr;
};

namespace evx {

int plan_create(evx_imex_plan** out, int nx, int ny, int nz, int is_f64, int backend);
int plan_destroy(evx_imex_plan* p);

// fft_generic.cu (mixed-radix passes: extents with prime factors <= 7, float32 / float64)
bool generic_fft_supported(int nx, int ny, int nz);
int generic_plan_init(evx_imex_plan* p);
template <typename R>
int generic_apply(evx_imex_plan* p, const R* u, const R* r, R* out, void* workspace,
                  const double* h, double dt, double coef, int power, cudaStream_t st);

// fft_native.cu
bool native_fft_supported(int nx, int ny, int nz);
int native_plan_init(evx_imex_plan* p);
void native_plan_free(evx_imex_plan* p);
int native_apply(evx_imex_plan* p, const float* u, const float* r, float* out, void* workspace,
                 const double* h, double dt, double coef, int power, cudaStream_t st);
int native_single_pass(evx_imex_plan* p, int which, const float* u, const float* r, float* out,
                       void* workspace, const double* h, double dt, double coef, int power,
                       cudaStream_t st);
int native_ch_step(evx_imex_plan* p, const float* u, const float* hom, float* out,
                   void* workspace, const double* h, double dt, double eps, double D, double A,
                   cudaStream_t st);

__device__ __forceinline__ bool aligned16_dev(const void* p) {
  return (reinterpret_cast<unsigned long long>(p) & 15ull) == 0;
}

}  // namespace evx
