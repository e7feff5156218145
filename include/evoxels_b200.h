/* evoxels_b200 - C ABI of the B200-native hot path of daubners/evoxels.
 *
 * One shared library (libevx_b200.so, built for sm_100a) exports everything below with
 * C linkage.  All data pointers are DEVICE pointers owned by the caller (PyTorch's caching
 * allocator in the shipped host code); no call allocates device memory except
 * evx_imex_plan_create (cuFFT plans, twiddle tables), no call synchronises the device,
 * every call enqueues its work on `stream` (a cudaStream_t passed as void*).  Return value:
 * 0 on success, > 0 a cudaError_t, < 0 one of EVX_ERR_* below; evx_strerror() names it.
 * Calls are thread-safe for distinct streams/buffers/plans.
 *
 * Fields are C-contiguous [nx, ny, nz] with z fastest - exactly the reference's
 * [C=1, Nx, Ny, Nz] tensors (evoxels/voxelgrid.py:126-130).  For x-slab decomposition `nx`
 * is the local slab thickness and the optional halo pointers carry the neighbour's planes.
 *
 * Each entry point names the reference interface (file:line in daubners/evoxels) it
 * replaces; INTEGRATION.md shows the binding a reference maintainer would add.
 */
#ifndef EVOXELS_B200_H
#define EVOXELS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVX_VERSION 100  /* 0.1.0 */

/* boundary-condition kinds per axis (evoxels/problem_definition.py:60-115) */
#define EVX_BC_PERIODIC 0
#define EVX_BC_NEUMANN 1
#define EVX_BC_DIRICHLET 2

/* error codes */
#define EVX_OK 0
#define EVX_ERR_ARG (-1)          /* null pointer / non-positive extent / bad enum      */
#define EVX_ERR_UNSUPPORTED (-2)  /* valid in the reference, not implemented on device  */
#define EVX_ERR_ALIGN (-3)        /* pointer not aligned as the entry point requires    */
#define EVX_ERR_CUFFT (-1000)     /* -1000 - cufftResult                                */

/* FFT back ends of an IMEX plan */
#define EVX_FFT_AUTO 0
#define EVX_FFT_CUFFT 1   /* cuFFT R2C/C2R + fused filter / add kernels (any extents)   */
#define EVX_FFT_NATIVE 2  /* hand-written sm_100a pass kernels (power-of-two extents)   */
#define EVX_FFT_NATIVE_MIXED 3 /* hand-written mixed-radix passes: any extents whose prime
                                  factors are <= 7 (e.g. the README's 100^3), fp32 and fp64 */

int evx_version(void);
const char* evx_strerror(int code);

/* ---------------------------------------------------------------------------------
 * Stencil layer
 * ------------------------------------------------------------------------------- */

/* Cahn-Hilliard right-hand side, fused: clip -> 7-pt Laplacian -> mu -> face-mobility
 * flux -> divergence, ghost layers by index arithmetic.
 * Replaces CahnHilliard.rhs (evoxels/problem_definition.py:328-371) incl. its calls to
 * FDStencils.laplace/to_*_face/grad_*_face (evoxels/fd_stencils.py:20-42,62-75) and
 * CellCenteredBCs.pad_* (evoxels/boundary_conditions.py:9-59).
 *   c        raw concentration (clipped to [0,1] inside, like the reference)
 *   hom      NULL -> default mu_hom 18/eps c(1-c)(1-2c); else a field holding mu_hom(clip(c))
 *   rhs      output, must not alias c
 *   h[3]     grid spacing;  bc_kind[3], bc_val[6]=(x_lo,x_hi,y_lo,y_hi,z_lo,z_hi) Dirichlet values
 *   halo_lo  NULL, or [2,ny,nz] raw planes x=-2,-1 of the x-neighbour slab (then the x rule
 *            of bc_kind is not applied on that side); halo_hi likewise planes x=nx,nx+1.
 */
int evx_ch_rhs_f32(const float* c, const float* hom, float* rhs, int nx, int ny, int nz,
                   const double* h, double eps, double D, const int* bc_kind,
                   const double* bc_val, const float* halo_lo, const float* halo_hi,
                   void* stream);
int evx_ch_rhs_f64(const double* c, const double* hom, double* rhs, int nx, int ny, int nz,
                   const double* h, double eps, double D, const int* bc_kind,
                   const double* bc_val, const double* halo_lo, const double* halo_hi,
                   void* stream);

/* Two-phase Allen-Cahn right-hand side k = rhs(phi), fused with the explicit-stage
 * arithmetic of ForwardEuler / RungeKutta4:
 *      k_out   = k                          (if k_out   != NULL)
 *      y_out   = base + alpha * k           (if y_out   != NULL; base may equal phi)
 *      acc_out = acc_in + beta * k          (if acc_out != NULL; acc_in NULL means 0)
 * Replaces TwoPhaseAllenCahn.rhs (evoxels/problem_definition.py:421-447) with
 * FDStencils.laplace / normal_laplace (evoxels/fd_stencils.py:44-103, 19-point footprint
 * incl. the |grad|^2 <= 1e-7 guard), the generic ghost rules of CellCenteredBCs.pad_bc
 * (evoxels/boundary_conditions.py:33-59, axis order x,y,z), ForwardEuler.step
 * (evoxels/timesteppers.py:42-43) and the axpys of RungeKutta4.step (:56-61).
 *   pot      NULL -> default potential 18/eps phi(1-phi)(1-2phi); else potential(clip(phi))
 *   halo_lo  NULL or [1,ny,nz] raw plane x=-1 of the neighbour slab; halo_hi plane x=nx.
 * Outputs must not alias phi (neighbouring threads read it).
 */
int evx_ac_stage_f32(const float* phi, const float* pot, float* k_out, const float* base,
                     float* y_out, double alpha, const float* acc_in, float* acc_out,
                     double beta, int nx, int ny, int nz, const double* h, double eps,
                     double gab, double M, double force, double curvature,
                     const int* bc_kind, const double* bc_val, const float* halo_lo,
                     const float* halo_hi, void* stream);
int evx_ac_stage_f64(const double* phi, const double* pot, double* k_out, const double* base,
                     double* y_out, double alpha, const double* acc_in, double* acc_out,
                     double beta, int nx, int ny, int nz, const double* h, double eps,
                     double gab, double M, double force, double curvature,
                     const int* bc_kind, const double* bc_val, const double* halo_lo,
                     const double* halo_hi, void* stream);

/* One ghost layer around a field: out[nx+2,ny+2,nz+2].
 * Replaces CellCenteredBCs.pad_periodic / pad_dirichlet_periodic / pad_zero_flux_periodic /
 * pad_bc (evoxels/boundary_conditions.py:9-59) and VoxelGridTorch.pad_periodic
 * (evoxels/voxelgrid.py:194-195); edge/corner ghosts follow the reference's x,y,z order. */
int evx_pad_ghost_f32(const float* in, float* out, int nx, int ny, int nz, const int* bc_kind,
                      const double* bc_val, void* stream);
int evx_pad_ghost_f64(const double* in, double* out, int nx, int ny, int nz, const int* bc_kind,
                      const double* bc_val, void* stream);

/* Stencils on an already ghost-padded field [nx+2,ny+2,nz+2] -> interior [nx,ny,nz]:
 * op 0 = 7-pt Laplacian (fd_stencils.py:62-75), 1 = normal Laplacian (:77-103),
 * 2 = |grad|^2 from centred differences (:56-60). */
int evx_padded_stencil_f32(const float* padded, float* out, int nx, int ny, int nz,
                           const double* h, int op, void* stream);
int evx_padded_stencil_f64(const double* padded, double* out, int nx, int ny, int nz,
                           const double* h, int op, void* stream);

/* Two-species reaction-diffusion rhs, u and out are [2,nx,ny,nz], fully periodic:
 *   out[0] = D_A lap7(u0) - I + feed (1 - u0),  out[1] = D_B lap7(u1) + I - kill u1,
 *   I = u0 u1^2, or the caller's `interaction` field [nx,ny,nz] when non-NULL.
 * Replaces CoupledReactionDiffusion.rhs (evoxels/problem_definition.py:614-633). */
int evx_rd2_rhs_f32(const float* u, const float* interaction, float* out, int nx, int ny, int nz,
                    const double* h, double D_A, double D_B, double feed, double kill,
                    void* stream);
int evx_rd2_rhs_f64(const double* u, const double* interaction, double* out, int nx, int ny,
                    int nz, const double* h, double D_A, double D_B, double feed, double kill,
                    void* stream);

/* ---------------------------------------------------------------------------------
 * Spectral (semi-implicit) stage
 * ------------------------------------------------------------------------------- */
typedef struct evx_imex_plan evx_imex_plan;

/* Plan for  out = u + irfftn( P(k) * rfftn(r) ),  P = dt / (1 + dt*coef*|k|^(2*power)),
 * on a periodic [nx,ny,nz] grid.  Replaces PseudoSpectralIMEX.__post_init__/step
 * (evoxels/timesteppers.py:75-89), VoxelGridTorch.rfftn/irfftn (evoxels/voxelgrid.py:
 * 203-207) and the stored prefactor array built from VoxelGrid.rfft_k_squared
 * (evoxels/voxelgrid.py:84-90,110-114): wavenumbers are recomputed on the fly in float32
 * with the reference's rounding sequence, nothing of size O(N^3) is stored.
 *   is_f64   0: float32 fields, 1: float64 fields (mixed-radix or cuFFT back end)
 *   backend  EVX_FFT_AUTO picks EVX_FFT_NATIVE for power-of-two float32 grids, else
 *            EVX_FFT_NATIVE_MIXED when every extent is 7-smooth, else EVX_FFT_CUFFT; asking
 *            for a back end that cannot do the grid returns EVX_ERR_UNSUPPORTED         */
int evx_imex_plan_create(evx_imex_plan** plan, int nx, int ny, int nz, int is_f64, int backend);
int evx_imex_plan_destroy(evx_imex_plan* plan);
int evx_imex_plan_backend(const evx_imex_plan* plan);   /* EVX_FFT_CUFFT|NATIVE|NATIVE_MIXED */
/* bytes of caller-provided scratch every apply/step call needs (256-byte aligned) */
int evx_imex_plan_workspace_bytes(const evx_imex_plan* plan, size_t* bytes);

/* out = u + irfftn( P * rfftn(r) ).  `r` is preserved, `out` may alias `u` but not `r`;
 * u == NULL gives the update alone (out = irfftn(P * rfftn(r))).
 * CH: coef = 2*eps*D*A, power = 2 (problem_definition.py:303).  AC / reaction-diffusion:
 * coef = M*gab or D*A, power = 1 (:198, :389).
 * OR-ing EVX_FILTER_ETD1 into `power` selects the exponential-Euler weight instead,
 *   P = dt * phi1(-dt*coef*|k|^(2*power)),  phi1(z) = (exp(z)-1)/z  (Pade branch |z|<0.5),
 * i.e. ExponentialEuler.step / phi1 / phiPade (evoxels/timesteppers.py:137-202).  The flag
 * is accepted wherever a `power` argument appears. */
#define EVX_FILTER_ETD1 0x100
/* Non-periodic x axis (zero-flux / Dirichlet; reference boundary_conditions.py:61-71 and
 * voxelgrid.py:116-124 mirror the field to 2 nx planes before rfftn): OR one of these into `power`
 * and pass the UN-extended [nx,ny,nz] arrays.  The y and z transforms then run on nx planes and the
 * x pass transforms each line together with its even / odd mirror image as one 2 nx-point line in
 * shared memory - the 2 nx array never exists in device memory.  Native back ends only (the
 * cuFFT back end answers EVX_ERR_UNSUPPORTED and the caller extends the field itself). */
#define EVX_FILTER_MIRROR_EVEN 0x200
#define EVX_FILTER_MIRROR_ODD 0x400
int evx_imex_apply_f32(evx_imex_plan* plan, const float* u, const float* r, float* out,
                       void* workspace, const double* h, double dt, double coef, int power,
                       void* stream);
int evx_imex_apply_f64(evx_imex_plan* plan, const double* u, const double* r, double* out,
                       void* workspace, const double* h, double dt, double coef, int power,
                       void* stream);

/* One full Cahn-Hilliard IMEX step on a fully periodic grid:
 *      out = u + irfftn( dt/(1 + dt*2*eps*D*A*|k|^4) * rfftn( CahnHilliard.rhs(u) ) )
 * = PseudoSpectralIMEX.step o CahnHilliard.rhs (timesteppers.py:85-89,
 * problem_definition.py:328-371).  `out` must not alias `u`. */
int evx_ch_imex_step_f32(evx_imex_plan* plan, const float* u, const float* hom, float* out,
                         void* workspace, const double* h, double dt, double eps, double D,
                         double A, void* stream);
int evx_ch_imex_step_f64(evx_imex_plan* plan, const double* u, const double* hom, double* out,
                         void* workspace, const double* h, double dt, double eps, double D,
                         double A, void* stream);

/* Measurement aid (bench.py per-kernel roofline): run ONE pass of the native pipeline on the
 * plan's scratch - which = 0 z forward (r -> spectrum), 1 y forward, 2 x forward*filter*inverse,
 * 3 y inverse, 4 z inverse (+u -> out); 5 = the chained z+y forward kernel, 6 = the chained y+z
 * inverse kernel (EVX_ERR_UNSUPPORTED when the plan runs one kernel per pass).  Native back
 * end only. */
int evx_imex_native_pass_f32(evx_imex_plan* plan, int which, const float* u, const float* r,
                             float* out, void* workspace, const double* h, double dt, double coef,
                             int power, void* stream);

/* The filter alone on a cuFFT-layout half spectrum [nx,ny,nz/2+1] (complex interleaved):
 * spec *= scale * dt / (1 + dt*coef*|k|^(2*power)).  Exposed for tests and for callers
 * that run their own transforms. */
int evx_spectral_filter_c64(void* spec, int nx, int ny, int nz, const double* h, double dt,
                            double coef, int power, double scale, void* stream);
int evx_spectral_filter_c128(void* spec, int nx, int ny, int nz, const double* h, double dt,
                             double coef, int power, double scale, void* stream);

/* ---------------------------------------------------------------------------------
 * x-slab distributed spectral stage (one process per GPU; rank owns x in [rank*nx/W, ...))
 *
 * The reference has no distributed code; this is the multi-GPU form of
 * PseudoSpectralIMEX.step (evoxels/timesteppers.py:85-89).  Power-of-two extents, W | nx, W | ny.
 * Buffers (all complex64, `*spec_bytes` bytes each, from evx_dist_plan_sizes):
 *   spec  [nx/W][ny][P]      local half spectrum (P = pitch, returned in *pitch)
 *   send / recv  [W][nx/W][ny/W][P]   all-to-all block layout: block j of `send` goes to rank j,
 *                block i of `recv` came from rank i, so `recv` is [nx][ny/W][P] (y-pencils).
 * Per step:  forward(r_local) -> all-to-all -> middle (x fwd * P(k)/N * x inv) -> all-to-all
 *            -> backward(u_local) = out_local.  The y passes read / write the block layout
 * directly, so there is no pack / unpack kernel around the collective.
 * ------------------------------------------------------------------------------- */
typedef struct evx_dist_plan evx_dist_plan;
int evx_dist_plan_create(evx_dist_plan** plan, int nx, int ny, int nz, int world, int rank);
int evx_dist_plan_destroy(evx_dist_plan* plan);
int evx_dist_plan_sizes(const evx_dist_plan* plan, size_t* spec_bytes, int* pitch);
/* cap the persistent grid of the peer-store launches (they are NVLink-bound; leaving SMs free
 * lets the next chunk's rhs / z pass run concurrently on another stream); 0 = fill the GPU */
int evx_dist_plan_set_p2p_ctas(evx_dist_plan* plan, int ctas);
int evx_dist_forward_f32(evx_dist_plan* plan, const float* r_local, void* spec, void* send,
                         void* stream);
int evx_dist_middle_f32(evx_dist_plan* plan, void* recv, const double* h, double dt, double coef,
                        int power, void* stream);
int evx_dist_backward_f32(evx_dist_plan* plan, const void* recv, void* spec, const float* u_local,
                          float* out_local, void* stream);
/* Fused transform + transpose over NVLink peer memory (W <= 8): instead of filling a local
 * send buffer for an all-to-all, the y pass (forward) / the x pass (middle) store every chunk
 * straight into the owning rank's buffer - peer_recv[j] / peer_out[j] is rank j's mapped
 * block buffer ([W][nx/W][ny/W][P], e.g. from torch symmetric memory); this rank fills block
 * `rank` of each.  The caller puts a cross-rank barrier on the stream afterwards. */
int evx_dist_forward_p2p_f32(evx_dist_plan* plan, const float* r_local, void* spec,
                             void* const* peer_recv, void* stream);
int evx_dist_middle_p2p_f32(evx_dist_plan* plan, void* recv, void* const* peer_out, const double* h,
                            double dt, double coef, int power, void* stream);
/* forward_p2p restricted to the local x planes [x0, x0+nxc) (pointers still address the full
 * local arrays); parts: 1 = z pass only, 2 = y pass (+peer stores) only, 3 = both.  Issuing the
 * z passes on a compute stream and the NVLink-bound y passes on a second stream overlaps the
 * transfer of one chunk with the rhs / z pass of the next. */
int evx_dist_forward_chunk_p2p_f32(evx_dist_plan* plan, const float* r_local, void* spec,
                                   void* const* peer_recv, int x0, int nxc, int parts, void* stream);
/* middle_p2p restricted to the local y-pencil rows [yl0, yl0+nylc).  The destination tables of the
 * *_chunk_p2p calls may mix peers' buffers with local ones (entry j - rank blocks before the local
 * buffer's block j): the hybrid transport stores the blocks of some ranks over NVLink from inside
 * the pass and leaves the others to the copy engines. */
int evx_dist_middle_chunk_p2p_f32(evx_dist_plan* plan, void* recv, void* const* peer_out, int yl0,
                                  int nylc, const double* h, double dt, double coef, int power,
                                  void* stream);

/* Copy-engine transport: the same block buffers, but the transposes are plain device-to-device
 * copies between mapped peer buffers issued on copy streams (DMA engines, no SM involved), so
 * that they overlap the kernels of the next chunk.  forward_chunk = z + y pass of the local x
 * planes [x0, x0+nxc) into block layout `send` (rows x0.. of every block are then contiguous:
 * nxc*(ny/W)*P elements per peer); middle_chunk = the x pass restricted to the local y-pencil
 * rows [yl0, yl0+nylc) of `recv` (a 2-D region per peer block: nx/W rows of nylc*P elements,
 * pitch (ny/W)*P).  A non-NULL `self_block` names the buffer that receives block `rank` (the
 * part that stays on this GPU) directly, so that no local copy is needed: the forward pass
 * then fills block `rank` of `self_block` (the local recv buffer) and the other blocks of
 * `send`; the middle pass writes rows x of block `rank` into `self_block` and the rest in place.
 * evx_copy_async / evx_copy2d_async enqueue cudaMemcpyAsync / cudaMemcpy2DAsync
 * (cudaMemcpyDefault) on `stream`. */
int evx_dist_forward_chunk_f32(evx_dist_plan* plan, const float* r_local, void* spec, void* send,
                               void* self_block, int x0, int nxc, void* stream);
/* n <= 8 copies of one chunk (region i: `height` rows of `width_bytes`, pitches dpitch / spitch;
 * height 1 = contiguous) handed to the driver as ONE batch (cudaMemcpyBatchAsync /
 * cudaMemcpy3DBatchAsync): the copies of a batch are unordered among themselves, so the driver may
 * spread them over its copy engines instead of running them one after the other. */
int evx_copy_batch_async(void* const* dst, size_t dpitch, const void* const* src, size_t spitch,
                         size_t width_bytes, size_t height, int n, void* stream);
int evx_dist_middle_chunk_f32(evx_dist_plan* plan, void* recv, void* self_block, int yl0, int nylc,
                              const double* h, double dt, double coef, int power, void* stream);
/* One launch that copies n <= 8 pitched regions (rows x row_bytes; 16-byte aligned) src[i] ->
 * dst[i] with `ctas_per_region` CTAs each: the SM-driven alternative to n DMA copies when the
 * regions are small and go to many peers (dst[i] = mapped peer memory -> NVLink stores). */
int evx_peer_scatter(const void* const* src, void* const* dst, int n, size_t row_bytes, size_t rows,
                     size_t src_pitch, size_t dst_pitch, int ctas_per_region, void* stream);
int evx_copy_async(void* dst, const void* src, size_t bytes, void* stream);
int evx_copy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes,
                     size_t height, void* stream);

/* ---------------------------------------------------------------------------------
 * Adjoint of the Cahn-Hilliard right-hand side (fully periodic grids)
 *
 * Backward pass of CahnHilliard.rhs / PseudoSpectralIMEX.step for parameter estimation -
 * the capability of evoxels/inversion.py:51-124 (JAX/diffrax there) behind a
 * torch.autograd.Function here (evoxels_b200/autograd.py).  With w = dL/d(rhs):
 *   evx_ch_mu            mu = (18/eps) p(c^) - 2 eps lap(c^),  c^ = clip(u,0,1)
 *   evx_ch_adjoint_flux  z = D div( c_f(1-c_f) grad w ),
 *                        m = -D/2 sum_faces (1 - 2 c_f) (dw)(dmu) / h^2
 *   evx_ch_adjoint_combine  lam_out = [lam_in +] 1[0<=u<=1] ( g'(c^) z - 2 eps lap(z) + m ),
 *                        *deps_acc += sum z ( -(18/eps^2) p(c^) - 2 lap(c^) )   (device double)
 * ------------------------------------------------------------------------------- */
int evx_ch_mu_f32(const float* u, float* mu, int nx, int ny, int nz, const double* h, double eps,
                  void* stream);
int evx_ch_mu_f64(const double* u, double* mu, int nx, int ny, int nz, const double* h, double eps,
                  void* stream);
int evx_ch_adjoint_flux_f32(const float* u, const float* mu, const float* w, float* z, float* m,
                            int nx, int ny, int nz, const double* h, double D, void* stream);
int evx_ch_adjoint_flux_f64(const double* u, const double* mu, const double* w, double* z,
                            double* m, int nx, int ny, int nz, const double* h, double D,
                            void* stream);
int evx_ch_adjoint_combine_f32(const float* u, const float* z, const float* m,
                               const float* lam_in, float* lam_out, double* deps_acc, int nx,
                               int ny, int nz, const double* h, double eps, void* stream);
int evx_ch_adjoint_combine_f64(const double* u, const double* z, const double* m,
                               const double* lam_in, double* lam_out, double* deps_acc, int nx,
                               int ny, int nz, const double* h, double eps, void* stream);
/* x-slab form (multi-GPU adjoint, SURVEY 8e last row): the arrays are the slab extended by halo
 * planes; lam_out is formed on every plane, *deps_acc only takes the planes [x_lo, x_hi) - the
 * slab's own - so that the partial sums of the ranks add up to the single-GPU value. */
int evx_ch_adjoint_combine_range_f32(const float* u, const float* z, const float* m,
                                     const float* lam_in, float* lam_out, double* deps_acc, int nx,
                                     int ny, int nz, const double* h, double eps, int x_lo, int x_hi,
                                     void* stream);
int evx_ch_adjoint_combine_range_f64(const double* u, const double* z, const double* m,
                                     const double* lam_in, double* lam_out, double* deps_acc, int nx,
                                     int ny, int nz, const double* h, double eps, int x_lo, int x_hi,
                                     void* stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
unsigned long long evx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EVOXELS_B200_H */
