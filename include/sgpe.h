/* sgpe.h — C ABI of the B200 split-step propagator (libsgpe.so).
 *
 * The reference (ultracoldYEG/spinor-gpe) is pure Python and has no FFI: its boundary for this path is
 * the class spinor_gpe/pspinor/tensor_propagator.py:TensorPropagator plus the helpers of
 * spinor_gpe/pspinor/tensor_tools.py.  Each entry point below names the reference code it replaces.
 * The Python binding a maintainer would add is shown in INTEGRATION.md (ctypes, ~40 lines).
 *
 * Conventions
 *  - every function returns 0 on success, a negative SGPE_E* code otherwise, never throws;
 *    sgpe_last_error() returns a thread-local message for the last failure;
 *  - pointers named *_dev are CUDA device pointers owned by the caller; *_host are host pointers;
 *  - a state is [batch][2][ny][nx] complex, x contiguous (the reference's list of two (Ny,Nx) arrays,
 *    tensor_propagator.py:118), complex128 (dtype 0) or complex64 (dtype 1), k-space states in the
 *    reference's fft-shifted order with its dx*dy/2pi scaling (tensor_tools.py:218-226);
 *  - operator grids (kinetic, potential, coupling) are real float64 [ny][nx] in both precisions, given
 *    per trajectory with a batch stride in elements (0 = shared by all trajectories);
 *  - all work is enqueued on the given stream; nothing synchronises the device except the *_host calls;
 *  - a plan is not thread-safe; distinct plans are independent.
 */
#ifndef SGPE_H
#define SGPE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgpe_plan sgpe_plan;
typedef void* sgpe_stream;            /* cudaStream_t */

enum { SGPE_C128 = 0, SGPE_C64 = 1 };
enum { SGPE_TIME_REAL = 0, SGPE_TIME_IMAG = 1 };
enum { SGPE_COUPLING_NONE = 0, SGPE_COUPLING_UNIFORM = 1, SGPE_COUPLING_DENSE = 2 };
enum {
    SGPE_OK = 0,
    SGPE_EINVAL = -1,      /* bad argument / unsupported size */
    SGPE_ECUDA = -2,       /* CUDA runtime error */
    SGPE_ESTATE = -3,      /* call order (e.g. stepping before operators are set) */
    SGPE_ENOMEM = -4
};

const char* sgpe_last_error(void);
/* library build info: "sgpe <version> sm_100a" (or "... emu" for the CPU test build) */
const char* sgpe_version(void);

/* nx, ny: powers of two in [32, 4096].  batch >= 1 independent trajectories.
 * Replaces the tensor set-up of TensorPropagator.__init__ (tensor_propagator.py:111-122). */
int sgpe_plan_create(sgpe_plan** out, int nx, int ny, int batch, int dtype, int device);
int sgpe_plan_destroy(sgpe_plan* p);

/* space['dr'], space['dv_r'], space['dv_k'], atom_num (tensor_propagator.py:111, 119-121). */
int sgpe_set_grid(sgpe_plan* p, double dx, double dy, double dv_r, double dv_k, double atom_num);
/* g_sc['uu'], ['dd'], ['ud'] (tensor_propagator.py:113, 247-248). */
int sgpe_set_interactions(sgpe_plan* p, double g_uu, double g_dd, double g_ud);
/* kin_eng_spin (tensor_propagator.py:114): borrowed device pointers, must outlive the stepping. */
int sgpe_set_kinetic(sgpe_plan* p, const double* kin0_dev, const double* kin1_dev, int64_t batch_stride);
/* pot_eng_spin (tensor_propagator.py:116). */
int sgpe_set_potential(sgpe_plan* p, const double* pot0_dev, const double* pot1_dev, int64_t batch_stride);
/* Separable fast path for the same two operators: kin_c[ky][kx] = kin_x[c][kx] + kin_y[c][ky] and
 * pot_c[y][x] = pot_x[c][x] + pot_y[c][y] (true for the reference's harmonic trap, free / Raman-shifted
 * dispersion, linear detuning gradients: pspinor.py:426-430, 496-501, 574-575).  kin_x_dev / pot_x_dev are
 * [2][nx], kin_y_dev / pot_y_dev [2][ny] (component-major); batch strides are 0 (shared) or 2*nx / 2*ny.
 * The library turns them into 1-D factor tables exp(-i e tau) per sub-step, so the passes multiply by
 * table products instead of evaluating exp / sincos per grid point.  The host decides whether the grids
 * are separable (spinor_gpe_b200.tensor_propagator does it to 1e-13 relative); the dense setters remain
 * the general path. */
int sgpe_set_kinetic_separable(sgpe_plan* p, const double* kin_x_dev, const double* kin_y_dev,
                               int64_t x_batch_stride, int64_t y_batch_stride);
int sgpe_set_potential_separable(sgpe_plan* p, const double* pot_x_dev, const double* pot_y_dev,
                                 int64_t x_batch_stride, int64_t y_batch_stride);
/* coupling / expon (tensor_propagator.py:122-129, tensor_tools.py:563-591).
 *   mode NONE    : is_coupling False (tensor_propagator.py:252, 258 skipped)
 *   mode UNIFORM : one Omega per trajectory, omega_dev[batch]
 *   mode DENSE   : coupling_dev [ny][nx] (+ batch stride)
 *   eiphi_dev    : exp(+i*expon) along x, nx complex values of the plan's dtype, or NULL when the
 *                  coupling is in the rotating frame (expon = 0, tensor_propagator.py:126-127). */
int sgpe_set_coupling(sgpe_plan* p, int mode, const double* coupling_dev, int64_t batch_stride,
                      const double* omega_dev, const void* eiphi_dev);
/* Coupling grid of the energy expectation when it is not the one the stepping applies: the reference's eng_expect
 * adds Re(conj(psi0) psi1) * coupling from `self.coupling` whatever `is_coupling` says (tensor_propagator.py:319-321),
 * while single_step applies the coupling operator only if is_coupling (:252, :258).  mode as in sgpe_set_coupling
 * (the Raman phase plays no role in the energy); -1 = follow sgpe_set_coupling (default). */
int sgpe_set_energy_coupling(sgpe_plan* p, int mode, const double* coupling_dev, int64_t batch_stride,
                             const double* omega_dev);
/* Tuning knobs.  "col_tile": 0 = default column-tile width (64-byte global segments, one CTA per SM at
 * 2048 points), 2 = half width (two CTAs per SM). */
int sgpe_set_option(sgpe_plan* p, const char* name, int value);
/* time = 'real' | 'imag' and t_step (tensor_propagator.py:96-103): fixes dt_out, dt_in. */
int sgpe_set_time(sgpe_plan* p, int time_mode, double dt);

/* Copy a k-space state into the plan (self.psik = to_tensor(spin.psik), tensor_propagator.py:118). */
int sgpe_load_psik(sgpe_plan* p, const void* psik_dev, sgpe_stream st);
/* Materialise the current normalised k-space state (what self.psik holds after a step,
 * tensor_propagator.py:271). */
int sgpe_store_psik(sgpe_plan* p, void* psik_dev, sgpe_stream st);

/* n x TensorPropagator.full_step() (tensor_propagator.py:214-222) with the per-step populations of
 * prop_loop (tensor_propagator.py:194): pops_dev (nullable) is [batch][pops_stride] doubles, step i
 * writes pops_dev[b*pops_stride + 2*(pops_first+i) + {0,1}].  Asynchronous. */
int sgpe_full_steps(sgpe_plan* p, int n, double* pops_dev, int64_t pops_stride, int pops_first,
                    sgpe_stream st);
/* The same with the energy expectation (eng_expect, tensor_propagator.py:273-324) of EVERY step boundary — the "energy
 * tracking" of BASELINE configs[2]; the reference evaluates the energy once per run, on the CPU.  energy_dev is
 * [batch][energy_stride] doubles, step i writes energy_dev[b*energy_stride + 4*(energy_first+i) + {0..3}] = E_total,
 * E_kin, E_pot, E_int.  The junction pass that follows a full step stores the boundary state on the side; its inverse
 * transform (with the normalisation and the density maxima folded into the last pass) and the stencil pass run
 * behind it: three extra launches per step, nothing synchronises — unwrap_mode 0 or 1 (see sgpe_energy).
 * unwrap_mode 2 (the reference's own definition: phase unwrapped by region merging) evaluates every closed step
 * through the stand-alone path of sgpe_energy and synchronises st once per step (≈ 30 ms per step at 2048^2). */
int sgpe_full_steps_energy(sgpe_plan* p, int n, double* pops_dev, int64_t pops_stride, int pops_first,
                           double* energy_dev, int64_t energy_stride, int energy_first, int unwrap_mode,
                           double kl_term, sgpe_stream st);
/* One TensorPropagator.single_step (tensor_propagator.py:224-271) of sub-step length dt_sub
 * (use sgpe_substeps for dt_out / dt_in). */
int sgpe_single_step(sgpe_plan* p, double dt_sub, sgpe_stream st);
int sgpe_substeps(const sgpe_plan* p, double* dt_out, double* dt_in);

/* ttools.fft_2d / ifft_2d (tensor_tools.py:201-258) and fft_1d / ifft_1d (:130-198; axis 0 = x,
 * 1 = y as in the reference) on a [batch][2][ny][nx] array, with the reference's shift and scaling
 * taken from sgpe_set_grid.  in == out is allowed. */
int sgpe_fft2d(sgpe_plan* p, const void* in_dev, void* out_dev, int inverse, sgpe_stream st);
int sgpe_fft1d(sgpe_plan* p, const void* in_dev, void* out_dev, int axis, int inverse, sgpe_stream st);

/* Per-component sum |psi_c|^2 (ttools.norm_sq / calc_pops without the volume element,
 * tensor_tools.py:437, 482): out_dev[batch][2]. */
int sgpe_sumsq(sgpe_plan* p, const void* in_dev, double* out_dev, sgpe_stream st);
/* ttools.norm (tensor_tools.py:289-305): out = in / sqrt(sum(|in|^2) * vol / atom_num). */
int sgpe_normalise(sgpe_plan* p, const void* in_dev, void* out_dev, double vol, sgpe_stream st);

/* TensorPropagator.eng_expect (tensor_propagator.py:273-324): [E_total, E_kin, E_pot, E_int] as raw grid
 * sums, out_dev[batch][4].  psik_dev == NULL evaluates the plan's current state.  kl_term is
 * 2*kL_recoil*is_coupling (:311).  unwrap_mode 0 leaves the wrapped phase as is (the variant pinned
 * against the reference), 1 differentiates with locally wrapped differences, 2 unwraps the
 * phase of each component the way the reference does (ttools.phase(psi, uwrap=True), tensor_propagator.py:304
 * -> skimage.restoration.unwrap_phase, tensor_tools.py:531; see sgpe_unwrap_phase) — mode 2 synchronises st. */
int sgpe_energy(sgpe_plan* p, const void* psik_dev, int unwrap_mode, double kl_term, double* out_dev,
                sgpe_stream st);

/* Spectral kinetic energy per component, out_dev[batch][2] = dv_k * sum_k kin_c(k) |psi_k,c(k)|^2  [hbar omega_x]:
 * the k-space counterpart of the finite-difference / unwrapped-phase kinetic term of eng_expect
 * (tensor_propagator.py:306-311), which needs no phase (no reference equivalent).  kin_c is the
 * plan's kinetic operator (kin_eng_spin: Raman shift and "- min" offset included, pspinor.py:496-501).
 * psik_dev == NULL evaluates the plan's current (normalised) state.  Asynchronous. */
int sgpe_kinetic_spectral(sgpe_plan* p, const void* psik_dev, double* out_dev, sgpe_stream st);

/* ttools.grad_comp (tensor_tools.py:331-350: np.gradient(psi_comp, *delta_r); the reference raises for tensors, :343-345)
 * on ONE (ny, nx) field of the plan's mesh that lives on the device: f_dev holds reals of the plan's precision
 * (is_complex = 0) or complex numbers (is_complex = 1: both parts are differentiated).  g0_dev = d/d(axis 0) with
 * spacing h0, g1_dev = d/d(axis 1) with spacing h1 — np.gradient's argument order, second-order central differences
 * inside, first-order one-sided at the edges.  Same element type and shape as f_dev; f must not alias g0 / g1. */
int sgpe_gradient(sgpe_plan* p, const void* f_dev, int is_complex, double h0, double h1, void* g0_dev, void* g1_dev,
                  sgpe_stream st);

/* The same functional on a REAL-space state [batch][2][ny][nx] of the plan's dtype (eng_expect after its ifft_2d,
 * tensor_propagator.py:301-324).  Valid on line plans too (sgpe_plan_create_lines: ny = nlines rows of nx = len points,
 * potential / coupling / interactions set as for sgpe_pass_rows), which is how meshes beyond 4096 points per line get
 * their energy (spinor_gpe_b200/slab.py on one device). */
int sgpe_energy_real_space(sgpe_plan* p, const void* psi_dev, int unwrap_mode, double kl_term, double* out_dev,
                           sgpe_stream st);

/* ttools.phase_comp(psi_comp, uwrap=True, dens) (tensor_tools.py:514-539): two-dimensional phase unwrapping by
 * reliability-sorted region merging (Herraez, Burton, Lalor, Gdeisat, Appl. Opt. 41, 7437 (2002)), the algorithm
 * behind skimage.restoration.unwrap_phase without mask and wrap-around (the reference's call, tensor_tools.py:531;
 * scikit-image==0.16.2, requirements.txt:25 — a third-party dependency restated from the publication).
 *   in_dev  : nplanes planes [ny][nx] (the plan's mesh); kind 0 = complex values of the plan's dtype (phase =
 *             atan2(im, re), np.angle at :529), kind 1 = float64 wrapped angles
 *   mask    : != 0 (kind 0 only) zeroes the result where |psi|^2 < 1e-6 * max |psi|^2 of the plane (:538)
 *   out_dev : nplanes x [ny][nx] float64 = wrapped phase + 2 pi * integer
 * Everything but a short tail runs on the device: angles, per-pixel reliabilities, edge keys, the radix sort of the
 * edges, the region merging (the minimum spanning tree of the pixel grid under the edge ranks with the relative
 * multiples of 2 pi along it — Boruvka rounds over an offset-carrying union-find — then the one pixel group whose
 * values the published merge rules never move, found level by level with a bisection over the rank threshold and a
 * lock-free, size-carrying union-find; groups of at most "unwrap_tail" = 16384 tree edges finish on the host) and the final pass.
 * The integer field equals the sequential edge-by-edge merging's bit for bit, global offset included.  Border pixels
 * get the fixed reliability value 9999999 (scikit-image adds rand()), equal keys keep edge order (horizontal edges
 * row by row, then vertical).  Synchronises st.
 * Options (sgpe_set_option), all for cross-checks: "unwrap_sort" = 1 sorts the edges on the host (same order);
 * "unwrap_merge" = 1 runs the whole merging on the host (sequential offset-carrying union-find over all edges, one
 * thread per plane); "unwrap_anchor" = 0 finds the surviving group by a sequential host pass over the tree edges
 * (default for more than 4 planes or fewer than 65536 pixels: one plane per core), 1 by the device bisection. */
int sgpe_unwrap_phase(sgpe_plan* p, const void* in_dev, int kind, int nplanes, int mask, double* out_dev,
                      sgpe_stream st);

/* ---- Slab-decomposed (multi-GPU) transforms: local building blocks.  The grid is split by rows over P
 * ranks; the caller owns the buffers and the collectives (all-to-all transpose, all-reduce of the norm
 * sums) — spinor_gpe_b200/slab.py does it with torch.distributed over NCCL.  The reference has no
 * counterpart (single device only, benchmarks/benchmark_prop.py:86-88 stops at out-of-memory).
 *  sgpe_pass_rows   : row_pass on a local [2][ny_local][nx] buffer of a plan(nx, ny_local): iFFT_x, norm with
 *                     the GLOBAL sum totals_dev[0] and global point count, I C P C I, FFT_x
 *                     (tensor_propagator.py:243-269).
 *  sgpe_pass_klines : the k-space junction on a transposed local buffer [2][nlines][len] of a
 *                     plan(nx = len, ny = nlines): [FFT] FA sums FB sums [iFFT] along the contiguous lines
 *                     (tensor_propagator.py:270-271 and :242 of the next sub-step); the plan's kinetic operator
 *                     is given line-major (dense [nlines][len], or separable kin_x <-> position, kin_y <-> line);
 *                     sums_dev[3] receives the local T, S0, S1.
 *  sgpe_slab_pack   : [2][lines][P*chunk] -> [P][2][lines][chunk]    (send buffer of the all-to-all)
 *  sgpe_slab_unpack : [P][2][h][w] -> [2][w][P*h], transposing every block (receive side). */
/* Line plans for the slab mode: `nlines` contiguous lines of `len` points, both components, caller-owned
 * buffers ([2][nlines][len]).  n1 == 1: len in [32, 4096].  n1 > 1: four-step split len = n1 * n2 for lines a
 * CTA cannot hold (16384 = 128 x 128): sgpe_pass_klines then transforms the contiguous sub-lines of n2 points
 * (its operator tables are indexed by the position n1-digit * n2 + n2-digit, i.e. k-space is kept in the
 * digit-transposed order  position k1*n2 + k2  <->  frequency k1 + n1*k2), and sgpe_pass_mid does the strided
 * half over n1 with the four-step twiddles and, optionally, the real-space operators fused in:
 *   [* conj w] [iFFT over k1] [norm, I C P C I] [FFT over n1] [* w]. */
int sgpe_plan_create_lines(sgpe_plan** out, int len, int nlines, int n1, int dtype, int device);
int sgpe_pass_mid(sgpe_plan* p, void* buf_dev, int pre_tw, int do_inv, int do_pw, double dt_sub, int do_fwd,
                  int post_tw, const double* totals_dev, double global_points, int inner, int scatter,
                  sgpe_stream st);
int sgpe_pass_rows(sgpe_plan* p, void* buf_dev, double dt_sub, const double* totals_dev, double global_points,
                   int scatter, sgpe_stream st);
int sgpe_pass_klines(sgpe_plan* p, void* buf_dev, int do_fwd, int has_a, double tau_a, int has_b, double tau_b,
                     int do_inv, double* sums_dev, int scatter, sgpe_stream st);
int sgpe_slab_pack(sgpe_plan* p, const void* in_dev, void* out_dev, int lines, int nranks, int chunk, sgpe_stream st);
int sgpe_slab_unpack(sgpe_plan* p, const void* in_dev, void* out_dev, int nranks, int block_h, int block_w,
                     sgpe_stream st);

/* ---- Fused exchange (compute + collective in ONE kernel over peer memory).  Instead of pack -> NCCL all-to-all ->
 * unpack, the last pass of each direction stores every element directly into the buffer of the rank that needs it
 * next (buffers of the other GPUs of the node mapped with CUDA IPC; the stores travel over NVLink / NVSwitch while
 * the CTA's neighbours are still transforming), so the transfer overlaps the math tile by tile and four HBM passes
 * per sub-step disappear.  Everything stays ROW-MAJOR (no transposes):
 *     row slab of rank r : [2][Ny/P][Nx]     k slab of rank q : [2][Ny][Nx/P]
 *  sgpe_slab_set_peers : destination map of a plan's scatter stores.  peer_bufs[q] = rank q's destination buffer as
 *                        seen from this device.  mode 1 (row direction -> k slabs): seg = Nx/P, drow = Nx/P,
 *                        dplane = Ny*Nx/P, base = this rank's first global row.  mode 2 (k direction -> row slabs):
 *                        seg = Ny/P, drow = Nx, dplane = (Ny/P)*Nx, base = this rank's first global column.
 *  scatter != 0 on a pass = "store through the map" (the input is still read from buf_dev).  Ordering between ranks is
 *                        the caller's: a collective (the all-reduce of the norm sums, or a barrier) after the storing
 *                        kernel orders it before the consumers on every rank.
 *  sgpe_pass_kcols     : the k-space junction on a k slab [2][len][nlines] of a line plan(len, nlines[, n1]): W adjacent
 *                        columns per CTA, [FFT] K_a sums K_b sums [iFFT] down the columns (with n1 > 1: over the n2
 *                        sub-lines of every n1 group, the strided halves being sgpe_pass_mid(inner = nlines)).
 *  sgpe_pass_mid(inner = nlines > 1) : the strided four-step half on the k slab viewed as [2][n1][n2][nlines].
 *  sgpe_ipc_*          : cudaMalloc'ed exchange buffers exported to / imported from the other processes of the node
 *                        (cudaIpcGetMemHandle / cudaIpcOpenMemHandle with lazy peer access).
 *  sgpe_slab_window    : chunked pipelining of the exchange.  The line passes that follow work on the lines
 *                        [first, first + count) of the plan only (rows of a row slab, columns of a k slab), with
 *                        reduction slot `chunk` (0..15; sums of a chunk go to the sums pointer of that call) and,
 *                        for the passes that walk their window with persistent CTAs (sgpe_pass_klines,
 *                        sgpe_pass_mid on the column slab), at most max_ctas CTAs (0: one CTA per block).  A
 *                        scatter pass capped to a few CTAs per SM on a second stream sends chunk c over NVLink
 *                        while the local passes of chunk c + 1 run beside it.  count == 0 clears the window. */
#define SGPE_IPC_HANDLE_BYTES 64
int sgpe_slab_window(sgpe_plan* p, int first, int count, int chunk, int max_ctas);
int sgpe_slab_set_peers(sgpe_plan* p, void* const* peer_bufs, int nranks, int mode, int seg, int drow, int64_t dplane,
                        int base);
int sgpe_pass_kcols(sgpe_plan* p, void* buf_dev, int do_fwd, int has_a, double tau_a, int has_b, double tau_b,
                    int do_inv, double* sums_dev, int scatter, sgpe_stream st);
int sgpe_ipc_alloc(int device, uint64_t bytes, void** dev_ptr, unsigned char handle[SGPE_IPC_HANDLE_BYTES]);
int sgpe_ipc_open(int device, const unsigned char handle[SGPE_IPC_HANDLE_BYTES], void** dev_ptr);
int sgpe_ipc_close(void* dev_ptr);
int sgpe_ipc_free(void* dev_ptr);

/* The same path with HOST buffers (pageable or pinned): H2D of the state, n full steps, D2H of the
 * final normalised state and the populations [batch][n][2]; synchronises the stream before returning.
 * Equivalent of PSpinor.imaginary()/real() minus file output (pspinor.py:912-925). */
int sgpe_run_host(sgpe_plan* p, const void* psik_in_host, void* psik_out_host, int n_steps,
                  double* pops_host, sgpe_stream st);

/* Traffic / launch accounting for one full step of the whole batch:
 *   algorithmic = 768 B (c128) or 384 B (c64) per grid point;
 *   actual      = bytes the kernels are designed to move (state + operator grids);
 *   launches    = kernel launches per full step in steady state. */
int sgpe_step_accounting(const sgpe_plan* p, uint64_t* algorithmic_bytes, uint64_t* actual_bytes,
                         int* launches);
/* Per-kernel timing with CUDA events on the launching stream: between begin and end every column /
 * row pass is bracketed by an event pair; end synchronises and returns the summed milliseconds and the
 * launch counts per kind (used by bench.py for the roofline of the dominant kernel). */
int sgpe_profile_begin(sgpe_plan* p);
int sgpe_profile_end(sgpe_plan* p, double* ms_col, uint64_t* n_col, double* ms_row, uint64_t* n_row);
/* Dev tool: when buf_dev != NULL every row-pass CTA (column-pass CTA after sgpe_set_option("timeline_kind", 1)) writes
 * 8 uint64 (globaltimer at 6 phase boundaries, -, SM id). */
int sgpe_debug_timeline(sgpe_plan* p, unsigned long long* buf_dev);
/* Number of kernels this plan has launched since creation. */
int sgpe_launch_count(const sgpe_plan* p, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* SGPE_H */
