// sgpe_api.cu — plan object and the C ABI declared in include/sgpe.h.
//
// State machine of a plan.  The working state buffer is always [B][2][ny][nx] complex and is in one of
//   KSPACE : a k-space state (reference order), possibly un-normalised; `scale_pending` says whether
//            totals[b][1..2] hold the sums needed to normalise it (ttools.norm, tensor_propagator.py:271);
//   MID    : the output of a row pass — (k_x, y) space — with the trailing kinetic half-step of the
//            sub-step `pending_dt` still to be applied by the next column pass.
// A single step is  col_pass(FFT_y, K_a(pending), sums, K_b(this), sums, iFFT_y) ; row_pass(...)  i.e.
// two HBM round trips; the junction is closed (col_pass with FFT_y, K_a only) when the k-space state is
// needed.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#ifndef SGPE_EMU
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#endif

#include "../../include/sgpe.h"
#include "kernels.cuh"
#include "unwrap.cuh"
#include "generic.cuh"
#include "launch.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define SGPE_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(SGPE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
    } while (0)

const double kGamma = 1.0 / (2.0 + std::cbrt(2.0));      // tensor_propagator.py:101

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

#define SGPE_MAX_CHUNKS 16
struct sgpe_plan {
    int nx = 0, ny = 0, batch = 0, dtype = 0, device = 0;
    size_t csize = 16;
    long long plane = 0;
    void* state = nullptr;          // working state, allocated on first use (sgpe_load_psik / sgpe_run_host): plans that only
                                    // serve transforms / reductions (tensor_tools helpers) never pay for it
    bool line_plan = false;         // sgpe_plan_create_lines: caller-owned buffers, no k-space state
    void* tw_x = nullptr;
    void* tw_y = nullptr;
    double* partials = nullptr;
    unsigned* counter = nullptr;
    double* totals = nullptr;       // propagator sums [B][4]
    double* totals_aux = nullptr;   // sumsq / normalise sums [B][4]
    double* pops_buf = nullptr; size_t pops_cap = 0;
    void* scratch = nullptr;        // state-sized work buffer (energy), allocated on first use
    double* maxdens = nullptr;
    int max_tiles = 0;
    // problem
    bool grid_set = false, g_set = false, kin_set = false, pot_set = false, time_set = false;
    double dx = 1, dy = 1, dv_r = 1, dv_k = 1, atom_num = 1;
    double g_uu = 0, g_dd = 0, g_ud = 0;
    const double* kin0 = nullptr; const double* kin1 = nullptr; long long kin_bs = 0;
    const double* pot0 = nullptr; const double* pot1 = nullptr; long long pot_bs = 0;
    // separable operators: kin_c[ky][kx] = kin_x[c][kx] + kin_y[c][ky], pot_c[y][x] = pot_x[c][x] + pot_y[c][y]
    int kin_mode = 0, pot_mode = 0;
    const double* kin_x = nullptr; const double* kin_y = nullptr; long long kin_xbs = 0, kin_ybs = 0;
    const double* pot_x = nullptr; const double* pot_y = nullptr; long long pot_xbs = 0, pot_ybs = 0;
    struct FactorTable { bool valid = false; int tm = 0; double tau = 0; void* x = nullptr; void* y = nullptr; uint64_t used = 0; };
    FactorTable kin_tab[6], pot_tab[4];
    uint64_t tab_clock = 0;
    unsigned* sm_slots = nullptr;
    unsigned long long* dbg = nullptr;   // dev tool (sgpe_debug_timeline)
    unsigned long long* dbg_col = nullptr; int dbg_kind = 0;   // option "timeline_kind": 0 row pass, 1 column pass
    int stagger_ns = 0;            // option "stagger_ns"
    int prefetch = 1;              // L2 prefetch of the next tile of each SM (option "prefetch")
    int row_mode = 0;              // 0: both components per thread, 1: split (one component per thread)
    // long lines (four-step): nx = n1 * n2 with the strided part n1 (1 = ordinary plan)
    int n1 = 1, n2 = 0; void* tw_mid = nullptr; void* tw4 = nullptr;
    int col_wsel = 0;              // column tile width selector (sgpe_set_option "col_tile")
    int col_kernel = 0;            // option "col_kernel": 0 = default choice, 1 = one tile per CTA, 2 / 3 = persistent + TMA staging
    struct TileMap { const void* ptr = nullptr; int w = 0; SgpeTileMap* map = nullptr; };
    std::vector<TileMap> tile_maps;   // column-tile descriptors of the buffers the persistent pass has run on
    int unwrap_sort = 0;           // edge sort of the phase unwrapping: 0 device radix sort, 1 host (option "unwrap_sort")
    int unwrap_merge = 0;          // region merging: 0 spanning tree on the device, 1 all on the host
    int unwrap_tail = 16384;       // device anchor search: groups of at most this many tree edges finish on the host (option "unwrap_tail")
    int unwrap_anchor = -1;        // pixel group that keeps its values: 0 host pass over the tree edges, 1 device bisection, -1 by plane count
    struct UnwrapCache {           // scratch of the phase unwrapping, kept between evaluations (unwrap_scratch)
        double* rel = nullptr; unsigned long long* keys = nullptr; unsigned long long* keys_sorted = nullptr;
        unsigned* vals = nullptr; unsigned* vals_sorted = nullptr; void* tmp = nullptr; size_t tmp_bytes = 0;
        long long* ctrl = nullptr;      // bisection state of the anchor search (unwrap.cuh)
        unsigned* counters = nullptr;   // [0] groups hooked in a Boruvka round, [1] tree edges selected, [8..15] anchor search
        uint32_t* order_host = nullptr; int32_t* inc_host = nullptr;
        int nplanes = 0; size_t plane = 0; bool device_sort = false;
    } unwrap;
    double* unwrap_phi = nullptr;  // [B][2][ny][nx] wrapped phases of sgpe_energy(unwrap_mode 2), on first use
    int* unwrap_inc = nullptr;     // [B][2][ny][nx] multiples of 2 pi (sgpe_energy with unwrap_mode 2), on first use
    // fused exchange of the slab mode (sgpe_slab_set_peers): where the scatter stores of this plan's passes go
    // window of the slab the next line passes work on (sgpe_slab_window): lines [first, first + count), reduction
    // slot `chunk`, cap on the grid of persistent launches; count == 0: the whole slab
    struct Window { int first = 0, count = 0, chunk = 0, max_ctas = 0; } win;
    struct Peers { void* ptr[SGPE_MAX_PEERS]; int n = 0, mode = 0, seg = 0, drow = 0, base = 0; long long dplane = 0; } peers;
    int cpl_mode = 0; const double* cpl = nullptr; long long cpl_bs = 0;
    const double* omega = nullptr; const void* eiphi = nullptr;
    // coupling grid of the energy expectation when it differs from what the stepping applies (sgpe_set_energy_coupling):
    // eng_expect always adds Re(conj(p0) p1) * coupling (tensor_propagator.py:319-321), the step only if is_coupling
    int ecpl_mode = -1; const double* ecpl = nullptr; long long ecpl_bs = 0; const double* eomega = nullptr;
    int tm = SGPE_TIME_IMAG; double dt = 0, dt_out = 0, dt_in = 0;
    // state machine
    enum Phase { EMPTY, KSPACE, MID } phase = EMPTY;
    double pending_dt = 0;
    bool scale_pending = false;
    double* pend_pops = nullptr; long long pend_stride = 0; int pend_slot = -1;
    // per-step energy tracking (sgpe_full_steps_energy): where the energy of the step boundary that the NEXT junction
    // pass materialises goes (slot < 0: none), and how it is evaluated
    double* pend_energy = nullptr; long long pend_estride = 0; int pend_eslot = -1;
    int track_unwrap = 0; double track_kl = 0.0;
    // meshes outside the fused kernels' lengths (any even size with prime factors 2, 3, 5, 7; generic.cuh): one kernel
    // per operation behind the same run_col / run_row calls
    bool generic = false;
    std::vector<int> rad_x, rad_y;
    uint64_t launches = 0;
    // CUDA-graph replay of the steady-state full step (option "graph"): six kernel nodes whose arguments do not change
    // from step to step (the populations slot comes from the device-side counter slot_ctr)
    int energy_kernel = 0;         // option "energy_kernel": 0 streaming (default), 1 tiled
    int energy_polar = 1;          // option "energy_polar": per-step tracking hands (|psi|, arg psi) to the stencil pass
    int use_graph = -1;            // -1: default choice (on), 0 off, 1 on
    uint64_t epoch = 0;            // bumped by every sgpe_set_* call: a captured graph is valid for one epoch
    int* slot_ctr = nullptr;       // [batch]
    int* eslot_ctr = nullptr;      // [batch] energy slot of the replayed step
    bool capturing = false;
#ifndef SGPE_EMU
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphExec_t graph_exec_e = nullptr;      // the same step with the energy side chain (sgpe_full_steps_energy)
    cudaStream_t cap_stream = nullptr;
    struct GraphKey { uint64_t epoch = 0; const void* pops = nullptr; long long stride = 0; int tm = -1; double dt = 0;
                      const void* energy = nullptr; long long estride = 0; int unwrap = 0; double kl = 0;
                      bool operator==(const GraphKey& o) const { return epoch == o.epoch && pops == o.pops && stride == o.stride && tm == o.tm && dt == o.dt &&
                                                                        energy == o.energy && estride == o.estride && unwrap == o.unwrap && kl == o.kl; } } graph_key, graph_key_e;
#endif
    // optional per-kernel timing (sgpe_profile_*): event pairs around column (kind 0) / row (kind 1) passes
    bool prof_on = false;
#ifndef SGPE_EMU
    struct ProfRec { int kind; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof;
#endif
};

namespace {

using sgpe::ColArgs;
using sgpe::RowArgs;

#ifndef SGPE_EMU
struct ProfScope {
    sgpe_plan* p; cudaStream_t st; int kind; cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(sgpe_plan* p_, int kind_, cudaStream_t st_) : p(p_), st(st_), kind(kind_) {
        if (p->prof_on) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    }
    ~ProfScope() {
        if (e0) { cudaEventRecord(e1, st); p->prof.push_back({kind, e0, e1}); }
    }
};
#else
struct ProfScope { ProfScope(sgpe_plan*, int, cudaStream_t) {} };
#endif


template <typename C>
void fill_scatter(const sgpe_plan* p, bool on, sgpe::Scatter<C>* sc) {
    memset(sc, 0, sizeof(*sc));
    if (!on) return;
    for (int i = 0; i < p->peers.n; i++) sc->peer[i] = static_cast<C*>(p->peers.ptr[i]);
    sc->mode = p->peers.mode; sc->seg = p->peers.seg; sc->drow = p->peers.drow; sc->base = p->peers.base;
    sc->dplane = p->peers.dplane;
}

// exp(-2 pi i num / den) with exact values on the axes
template <typename T>
typename sgpe::cx_of<T>::type unit_root(long long num, long long den) {
    typename sgpe::cx_of<T>::type w;
    num %= den;
    const long double two_pi = 6.283185307179586476925286766559L;
    long double ang = two_pi * (long double)num / (long double)den;
    long double c = cosl(ang), s = sinl(ang);
    if (num == 0) { c = 1; s = 0; }
    if (4 * num == den) { c = 0; s = 1; }
    if (2 * num == den) { c = -1; s = 0; }
    if (4 * num == 3 * den) { c = 0; s = -1; }
    w.x = (T)c; w.y = (T)(-s);
    return w;
}

// Twiddle tables of a length-n transform for the two per-thread radices the kernels use, [E=8][E=16], n
// entries each.  For the plan with E elements per thread the Stockham stage whose previous radices
// multiply to Ns (Ns = E, E^2, ...) has radix R = min(E, n/Ns) and needs w_{Ns R}^{t k}, t = 1..R-1,
// k = 0..Ns-1; these are stored [t-1][k] (k fastest) starting at entry Ns - E (sum_{s'<s}(R-1)Ns' telescopes).
template <typename T>
int upload_twiddles(void** dst, int n) {
    typedef typename sgpe::cx_of<T>::type C;
    std::vector<C> h(2 * (size_t)n);
    for (auto& z : h) { z.x = (T)1; z.y = (T)0; }
    const int radices[2] = {8, 16};
    for (int e = 0; e < 2; e++) {
        const int E = radices[e];
        C* tab = h.data() + (size_t)e * n;
        for (long long Ns = E; Ns < n; Ns *= E) {
            const long long R = (n / Ns) < E ? (n / Ns) : E;
            for (long long t = 1; t < R; t++)
                for (long long k = 0; k < Ns; k++) tab[(Ns - E) + (t - 1) * Ns + k] = unit_root<T>(t * k, Ns * R);
        }
    }
    SGPE_CUDA(cudaMalloc(dst, sizeof(C) * h.size()));
    SGPE_CUDA(cudaMemcpy(*dst, h.data(), sizeof(C) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

// tau for exp(-i E tau): real time (len, 0), imaginary time (0, -len)
void time_arg(int tm, double len, double* re, double* im) {
    if (tm == SGPE_TIME_REAL) { *re = len; *im = 0.0; } else { *re = 0.0; *im = -len; }
}

// factor tables exp(-i * e * tau) of a separable operator for time argument tau (cached per plan)
template <typename T>
int factor_table(sgpe_plan* p, sgpe_plan::FactorTable* slots, int nslots, const double* ex, long long xbs,
                 const double* ey, long long ybs, double tau, cudaStream_t st, sgpe_plan::FactorTable** out) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe_plan::FactorTable* pick = nullptr;
    for (int i = 0; i < nslots; i++)
        if (slots[i].valid && slots[i].tm == p->tm && slots[i].tau == tau) { pick = &slots[i]; break; }
    if (!pick) {
        pick = &slots[0];
        for (int i = 0; i < nslots; i++) {
            if (!slots[i].valid) { pick = &slots[i]; break; }
            if (slots[i].used < pick->used) pick = &slots[i];
        }
        const long long nxe = (xbs ? xbs * p->batch : 2LL * p->nx), nye = (ybs ? ybs * p->batch : 2LL * p->ny);
        if (!pick->x) {
            SGPE_CUDA(cudaMalloc(&pick->x, sizeof(C) * (size_t)(2LL * p->nx * p->batch)));
            SGPE_CUDA(cudaMalloc(&pick->y, sizeof(C) * (size_t)(2LL * p->ny * p->batch)));
        }
        sgpe::ExpTableArgs<T> a;
        a.tm = p->tm;
        time_arg(p->tm, tau, &a.tr, &a.ti);
        a.e = ex; a.out = static_cast<C*>(pick->x); a.n = nxe;
        SGPE_LAUNCH((sgpe::exp_table<T>), dim3((unsigned)((nxe + 255) / 256)), dim3(256), 0, st, a);
        a.e = ey; a.out = static_cast<C*>(pick->y); a.n = nye;
        SGPE_LAUNCH((sgpe::exp_table<T>), dim3((unsigned)((nye + 255) / 256)), dim3(256), 0, st, a);
        p->launches += 2;
        SGPE_CUDA(cudaGetLastError());
        pick->valid = true; pick->tm = p->tm; pick->tau = tau;
    }
    pick->used = ++p->tab_clock;
    *out = pick;
    return 0;
}

void invalidate_tables(sgpe_plan::FactorTable* slots, int n) { for (int i = 0; i < n; i++) slots[i].valid = false; }

// Which column-pass kernel a plan uses unless sgpe_set_option("col_kernel") says otherwise (kernel_sel of ColArgs):
// measured on B200 (profiles/r02_kernel_sweep.jsonl) — complex128: the persistent pass with TMA-staged tiles and the
// split inverse exchange wins from 1024-point columns on (+5 % at 2048, +13 % at 4096; +12..17 % in real time) and for
// batched plans (+6..8 % at 8 / 64 x 512^2); complex64: half-width persistent tiles, two CTAs per SM (+12..19 % at 2048).
// With the plan's 32 KiB twiddle tables copied to shared memory (selector 4) the imaginary-time pass of ONE 2048-point
// trajectory gains another 0.7 % (109.9 -> 108.6 us); in real time, at 1024 points and for batches it loses 1.6-5 %, at
// 4096 points the tables do not fit beside the tile images (profiles/r02_variants.md).
int default_col_kernel(const sgpe_plan* p) {
    if (p->dtype == SGPE_C128) {
        if (p->ny == 2048 && p->batch == 1 && p->tm == SGPE_TIME_IMAG && p->kin_mode == 1) return 4;   // (factor tables only)
        return (p->ny >= 1024 || (p->batch >= 2 && p->ny >= 512)) ? 1 : 0;
    }
    return p->ny >= 1024 ? 5 : 0;
}

// Descriptor of the column tiles of a state buffer [B][2][ny][nx] for the persistent column pass: the buffer seen as a
// 2-D array of B*2*ny rows of 2*nx reals, boxes of 256 rows x one tile width (64 bytes).
int tile_map_for(sgpe_plan* p, const void* buf, int w, const SgpeTileMap** out) {
    for (auto& t : p->tile_maps)
        if (t.ptr == buf && t.w == w) { *out = t.map; return 0; }
    SgpeTileMap* m = new SgpeTileMap();
#ifndef SGPE_EMU
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            cudaGetLastError();
            delete m;
            return fail(SGPE_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        }
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {2ull * (cuuint64_t)p->nx, (cuuint64_t)p->batch * 2ull * (cuuint64_t)p->ny};
    const cuuint64_t strides[1] = {(cuuint64_t)p->nx * p->csize};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * w), (cuuint32_t)(p->ny < 256 ? p->ny : 256)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&m->m, p->dtype == SGPE_C128 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<void*>(buf), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        delete m;
        return fail(SGPE_ECUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
    }
#else
    (void)w;
    m->base = static_cast<const unsigned char*>(buf);
    m->row_stride = (long long)p->nx * (long long)p->csize;
    m->elem_bytes = (int)p->csize / 2;
#endif
    if (p->tile_maps.size() >= 8) { delete p->tile_maps.front().map; p->tile_maps.erase(p->tile_maps.begin()); }
    p->tile_maps.push_back({buf, w, m});
    *out = m;
    return 0;
}

// ---- generic meshes (generic.cuh)
// stage radices of a length-n transform: 4s first, then 2, 3, 5, 7; empty when n has another prime factor
std::vector<int> generic_radices(int n) {
    std::vector<int> r;
    while (n % 4 == 0) { r.push_back(4); n /= 4; }
    for (int q : {2, 3, 5, 7})
        while (n % q == 0) { r.push_back(q); n /= q; }
    if (n != 1 || r.size() > SGPE_GEN_MAX_STAGES) r.clear();
    return r;
}
bool generic_length(int n) { return n >= 2 && n % 2 == 0 && n <= 4096 && !generic_radices(n).empty(); }

template <typename T>
int gen_fft(sgpe_plan* p, const void* in, void* out, int axis /*0: along x, 1: along y*/, int dir, cudaStream_t st) {
    sgpe::GenFftArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.out = out;
    const std::vector<int>& rad = axis == 0 ? p->rad_x : p->rad_y;
    a.n = axis == 0 ? p->nx : p->ny;
    a.nlines = axis == 0 ? p->ny : p->nx;
    a.elem_stride = axis == 0 ? 1 : p->nx;
    a.line_stride = axis == 0 ? p->nx : 1;
    a.plane = p->plane; a.nplanes = 2 * p->batch; a.dir = dir;
    a.nstages = (int)rad.size();
    for (int i = 0; i < a.nstages; i++) a.radix[i] = rad[i];
    long long lpc = (200LL * 1024) / (2LL * a.n * (long long)p->csize);
    if (lpc < 1) lpc = 1;
    if (lpc > 16) lpc = 16;
    if (lpc > a.nlines) lpc = a.nlines;
    a.lpc = (int)lpc;
    const size_t smem = 2 * (size_t)a.lpc * a.n * p->csize;
#ifndef SGPE_EMU
    static size_t allowed = 0;
    if (smem > allowed) {
        SGPE_CUDA(cudaFuncSetAttribute(sgpe::gen_fft_pass<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(208 * 1024)));
        allowed = 208 * 1024;
    }
#endif
    const int groups = (a.nlines + a.lpc - 1) / a.lpc;
    SGPE_LAUNCH((sgpe::gen_fft_pass<T>), dim3((unsigned)(groups * a.nplanes)), dim3(256), smem, st, a);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int gen_scale(sgpe_plan* p, const void* in, void* out, int sign_x, int sign_y, double scale, const double* scale_tot,
              double scale_num, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::GenScaleArgs<T> a;
    a.in = static_cast<const C*>(in); a.out = static_cast<C*>(out); a.nx = p->nx;
    a.total = (long long)p->batch * 2 * p->plane; a.sign_x = sign_x; a.sign_y = sign_y; a.scale = scale;
    a.scale_tot = scale_tot; a.scale_num = scale_num; a.per_batch = 2 * p->plane;
    long long blocks = (a.total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    SGPE_LAUNCH((sgpe::gen_scale_sign<T>), dim3((unsigned)blocks), dim3(256), 0, st, a);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

int gen_blocks(const sgpe_plan* p) {
    long long b = (p->plane + 255) / 256;
    if (b > 256) b = 256;                       // four partial sums per CTA fit the reduction scratch
    if (b < 1) b = 1;
    return (int)b;
}

// the column pass of a generic mesh: [sign] -> [FFT_y] -> [factors + sums] -> [iFFT_y] -> [sign / scale]
template <typename T>
int run_col_generic(sgpe_plan* p, ColArgs<T>& a, bool fwd, bool inv, cudaStream_t st) {
    const void* cur = a.in;
    void* out = a.out;
    int rc;
    if (a.sign_in) { if ((rc = gen_scale<T>(p, cur, out, 0, 1, 1.0, nullptr, 0.0, st))) return rc; cur = out; }
    if (fwd) { if ((rc = gen_fft<T>(p, cur, out, 1, -1, st))) return rc; cur = out; }
    if (a.has_a || a.has_b) {
        a.in = static_cast<const typename sgpe::cx_of<T>::type*>(cur);
        dim3 grid((unsigned)gen_blocks(p), p->batch), block(256);
        if (p->tm == SGPE_TIME_REAL) { SGPE_LAUNCH((sgpe::gen_kspace_pass<T, sgpe::TM_REAL>), grid, block, 32 * 4 * sizeof(double), st, a); }
        else { SGPE_LAUNCH((sgpe::gen_kspace_pass<T, sgpe::TM_IMAG>), grid, block, 32 * 4 * sizeof(double), st, a); }
        p->launches++;
        SGPE_CUDA(cudaGetLastError());
        cur = out;
    }
    if (inv) { if ((rc = gen_fft<T>(p, cur, out, 1, +1, st))) return rc; cur = out; }
    if (a.sign_out || a.scale_out != 1.0 || cur != out) {
        if ((rc = gen_scale<T>(p, cur, out, 0, a.sign_out ? 1 : 0, a.scale_out, nullptr, 0.0, st))) return rc;
    }
    return 0;
}

// the row pass of a generic mesh: [sign] -> [iFFT_x] -> [normalise, I C P C I] -> [FFT_x] -> [sign / scale]
template <typename T>
int run_row_generic(sgpe_plan* p, RowArgs<T>& a, bool inv, bool pw, bool fwd, cudaStream_t st) {
    if (a.sc.mode || a.maxbits != nullptr) return fail(SGPE_EINVAL, "not available on generic meshes");
    const void* cur = a.in;
    void* out = a.out;
    int rc;
    if (a.sign_in) { if ((rc = gen_scale<T>(p, cur, out, a.sign_in & 1, (a.sign_in >> 1) & 1, 1.0, nullptr, 0.0, st))) return rc; cur = out; }
    if (inv) { if ((rc = gen_fft<T>(p, cur, out, 0, +1, st))) return rc; cur = out; }
    if (pw) {
        a.in = static_cast<const typename sgpe::cx_of<T>::type*>(cur);
        long long blocks = (p->plane + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        dim3 grid((unsigned)blocks, p->batch), block(256);
        if (p->tm == SGPE_TIME_REAL) { SGPE_LAUNCH((sgpe::gen_rspace_pass<T, sgpe::TM_REAL>), grid, block, 0, st, a); }
        else { SGPE_LAUNCH((sgpe::gen_rspace_pass<T, sgpe::TM_IMAG>), grid, block, 0, st, a); }
        p->launches++;
        SGPE_CUDA(cudaGetLastError());
        cur = out;
    }
    if (fwd) { if ((rc = gen_fft<T>(p, cur, out, 0, -1, st))) return rc; cur = out; }
    if (a.sign_out || a.scale_out != 1.0 || a.scale_tot != nullptr || cur != out) {
        if ((rc = gen_scale<T>(p, cur, out, a.sign_out & 1, (a.sign_out >> 1) & 1, a.scale_out, a.scale_tot, a.scale_num, st))) return rc;
    }
    return 0;
}

// Column pass.  tau_a / tau_b: time arguments of the k-space factors FA / FB (exp(-i kin tau)); has_a / has_b
// select them.
template <typename T>
int run_col(sgpe_plan* p, const void* in, void* out, bool fwd, bool has_a, double tau_a, bool has_b, double tau_b,
            bool inv, int sign_in, int sign_out, double scale_out, double* pops, long long pops_stride,
            int pops_slot, cudaStream_t st, void* aux = nullptr, double* zero2 = nullptr) {
    typedef typename sgpe::cx_of<T>::type C;
    ColArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.in = static_cast<const C*>(in); a.out = static_cast<C*>(out);
    a.tw = static_cast<const C*>(p->tw_y);
    a.nx = p->nx; a.ny = p->ny; a.plane = p->plane;
    a.do_fwd = fwd; a.do_inv = inv; a.has_a = has_a; a.has_b = has_b;
    a.prefetch_ahead = p->prefetch;
    a.sign_in = sign_in; a.sign_out = sign_out; a.scale_out = scale_out;
    a.kin_mode = p->kin_mode;
    if (has_a || has_b) {
        if (p->kin_mode == 0) {
            a.kin0 = p->kin0; a.kin1 = p->kin1; a.kin_bstride = p->kin_bs;
            time_arg(p->tm, tau_a, &a.ka_re, &a.ka_im);
            time_arg(p->tm, tau_b, &a.kb_re, &a.kb_im);
        } else {
            sgpe_plan::FactorTable* t = nullptr;
            int rc;
            if (has_a) {
                if ((rc = factor_table<T>(p, p->kin_tab, 6, p->kin_x, p->kin_xbs, p->kin_y, p->kin_ybs, tau_a, st, &t))) return rc;
                a.xa = static_cast<const C*>(t->x); a.ya = static_cast<const C*>(t->y);
            }
            if (has_b) {
                if ((rc = factor_table<T>(p, p->kin_tab, 6, p->kin_x, p->kin_xbs, p->kin_y, p->kin_ybs, tau_b, st, &t))) return rc;
                a.xb = static_cast<const C*>(t->x); a.yb = static_cast<const C*>(t->y);
            }
            a.sepx_bstride = p->kin_xbs ? p->kin_xbs : 0; a.sepy_bstride = p->kin_ybs ? p->kin_ybs : 0;
        }
    }
    a.partials = p->partials; a.counter = p->counter; a.totals = p->totals;
    a.pops = pops; a.pops_bstride = pops_stride; a.pops_slot = pops_slot;
    a.slot_ctr = (p->capturing && pops != nullptr && pops_slot >= 0) ? p->slot_ctr : nullptr;
    a.atom_num = p->atom_num;
    a.aux = static_cast<C*>(aux);
    a.zero2 = reinterpret_cast<unsigned long long*>(zero2);
    if (p->generic) return run_col_generic<T>(p, a, fwd, inv, st);
    a.dbg = (fwd && inv) ? p->dbg_col : nullptr;
    a.kernel_sel = p->col_kernel == 0 ? default_col_kernel(p) : p->col_kernel - 1;
    const bool inv_only = !fwd && inv && !has_a && !has_b && !sign_in && !sign_out && scale_out == 1.0;
    if (a.kernel_sel >= 1 && (fwd || inv_only) && inv && in == out && p->n1 == 1) {
        const SgpeTileMap* tm = nullptr;
        int w = sgpe::col_tile_width(p->ny, p->dtype);
        if (a.kernel_sel == 5) w /= 2;                      // half-width tiles, two CTAs per SM
        int rcm = tile_map_for(p, in, w, &tm);
        if (rcm) return rcm;
        a.tile_map = tm;
    }
    ProfScope prof(p, 0, st);
    int rc = sgpe::launch_col(p->ny, p->dtype, p->tm, &a, p->batch, p->col_wsel, st);
    if (rc == -3) return fail(SGPE_EINVAL, "column pass variant not compiled in (build with -DSGPE_EXPERIMENTAL)");
    if (rc != 0) return fail(SGPE_EINVAL, "column pass: unsupported geometry");
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int run_row(sgpe_plan* p, const void* in, void* out, bool inv, bool pw, double dt_sub, bool fwd, int sign_in,
            int sign_out, double scale_out, cudaStream_t st, const double* totals_override = nullptr,
            double norm_points = 0.0, bool scatter = false, const double* scale_tot = nullptr, double scale_num = 0.0,
            double* maxdens = nullptr, bool polar = false) {
    typedef typename sgpe::cx_of<T>::type C;
    RowArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.in = static_cast<const C*>(in); a.out = static_cast<C*>(out);
    a.tw = static_cast<const C*>(p->tw_x);
    a.nx = p->nx; a.ny = p->ny; a.plane = p->plane;
    a.do_inv = inv; a.do_pw = pw; a.do_fwd = fwd;
    a.prefetch_ahead = p->prefetch; a.stagger_ns = p->stagger_ns; a.dbg = pw ? p->dbg : nullptr;
    if (p->stagger_ns > 0) {
        if (!p->sm_slots) SGPE_CUDA(cudaMalloc((void**)&p->sm_slots, 1024 * sizeof(unsigned)));
        SGPE_CUDA(cudaMemsetAsync(p->sm_slots, 0, 1024 * sizeof(unsigned), st));
        a.sm_slots = p->sm_slots;
    }
    a.sign_in = sign_in; a.sign_out = sign_out; a.scale_out = scale_out;
    a.pot0 = p->pot0; a.pot1 = p->pot1; a.pot_bstride = p->pot_bs;
    a.pot_mode = pw ? p->pot_mode : 0;
    if (pw && p->pot_mode == 1) {
        sgpe_plan::FactorTable* t = nullptr;
        int rc0 = factor_table<T>(p, p->pot_tab, 4, p->pot_x, p->pot_xbs, p->pot_y, p->pot_ybs, dt_sub, st, &t);
        if (rc0) return rc0;
        a.px = static_cast<const C*>(t->x); a.py = static_cast<const C*>(t->y);
        a.sepx_bstride = p->pot_xbs; a.sepy_bstride = p->pot_ybs;
    }
    a.cpl_mode = p->cpl_mode; a.coupling = p->cpl; a.cpl_bstride = p->cpl_bs; a.omega_b = p->omega;
    a.eiphi = static_cast<const C*>(p->eiphi);
    a.g_uu = p->g_uu; a.g_dd = p->g_dd; a.g_ud = p->g_ud;
    time_arg(p->tm, dt_sub / 2, &a.ti_re, &a.ti_im);        // evolution_op(t_step / 2, int_eng), :249
    time_arg(p->tm, dt_sub, &a.tp_re, &a.tp_im);            // evolution_op(dt, pot_eng_spin), :140, 146
    a.tc = dt_sub / 4;                                      // coupling_op: Omega * dt_sub / 4, :142-149
    a.totals = totals_override ? totals_override : p->totals;
    a.norm_c = p->atom_num / (p->dv_r * (norm_points > 0 ? norm_points : (double)p->nx * (double)p->ny));
    fill_scatter(p, scatter, &a.sc);
    a.scale_tot = scale_tot; a.scale_num = scale_num;
    a.maxbits = reinterpret_cast<unsigned long long*>(maxdens);
    a.polar = polar ? 1 : 0;
    if (p->generic) return run_row_generic<T>(p, a, inv, pw, fwd, st);
    ProfScope prof(p, 1, st);
    int rc = sgpe::launch_row(p->nx, p->dtype, p->tm, &a, p->batch, p->row_mode, st);
    if (rc == -3) return fail(SGPE_EINVAL, "row pass variant not compiled in (build with -DSGPE_EXPERIMENTAL)");
    if (rc != 0) return fail(SGPE_EINVAL, "row pass: unsupported geometry");
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

#define SGPE_BY_DTYPE(p, fn, ...) ((p)->dtype == SGPE_C128 ? fn<double>(__VA_ARGS__) : fn<float>(__VA_ARGS__))

int ensure_state(sgpe_plan* p) {
    if (p->state) return 0;
    if (p->line_plan) return fail(SGPE_ESTATE, "line plans work on caller-owned buffers");
    if (cudaMalloc(&p->state, (size_t)p->batch * 2 * p->plane * p->csize) != cudaSuccess) {
        cudaGetLastError();
        return fail(SGPE_ENOMEM, "device allocation of the working state failed");
    }
    return 0;
}

int ready_to_step(sgpe_plan* p) {
    if (!p) return fail(SGPE_EINVAL, "null plan");
    if (!(p->grid_set && p->g_set && p->kin_set && p->pot_set && p->time_set))
        return fail(SGPE_ESTATE, "set grid, interactions, kinetic, potential and time before stepping");
    if (p->phase == sgpe_plan::EMPTY) return fail(SGPE_ESTATE, "no state loaded (sgpe_load_psik)");
    return 0;
}

int energy_of_boundary(sgpe_plan* p, cudaStream_t st);

int single_step_impl(sgpe_plan* p, double dt_sub, cudaStream_t st) {
    const bool mid = (p->phase == sgpe_plan::MID);
    const bool want_s = mid && p->pend_pops != nullptr && p->pend_slot >= 0;
    const bool want_e = mid && p->pend_energy != nullptr && p->pend_eslot >= 0;
    int rc;
    if (want_e) {
        // energy tracking: the junction also stores the boundary state (after the trailing half-step) on the side;
        // that needs the two factors separately in real time as well
        rc = SGPE_BY_DTYPE(p, run_col, p, p->state, p->state, true, true, p->pending_dt / 2, true, dt_sub / 2, true,
                           0, 0, 1.0, want_s ? p->pend_pops : nullptr, p->pend_stride, want_s ? p->pend_slot : -1, st,
                           p->scratch);
        if (!rc) rc = energy_of_boundary(p, st);
    } else if (!mid) {
        // leading kinetic half-step only (tensor_propagator.py:242)
        rc = SGPE_BY_DTYPE(p, run_col, p, p->state, p->state, false, false, 0.0, true, dt_sub / 2, true, 0, 0, 1.0,
                           nullptr, 0, -1, st);
    } else if (want_s && p->tm == SGPE_TIME_IMAG) {
        // the populations need the norms after the trailing half-step: two separate factors
        rc = SGPE_BY_DTYPE(p, run_col, p, p->state, p->state, true, true, p->pending_dt / 2, true, dt_sub / 2, true,
                           0, 0, 1.0, p->pend_pops, p->pend_stride, p->pend_slot, st);
    } else {
        // trailing half-step of the previous sub-step and leading one of this sub-step as ONE factor
        // exp(-i kin (dt_a + dt_b)/2); in real time |K| = 1 so the population sums are unaffected
        rc = SGPE_BY_DTYPE(p, run_col, p, p->state, p->state, true, false, 0.0, true, (p->pending_dt + dt_sub) / 2,
                           true, 0, 0, 1.0, want_s ? p->pend_pops : nullptr, p->pend_stride,
                           want_s ? p->pend_slot : -1, st);
    }
    if (rc) return rc;
    if (mid) p->pend_slot = -1;
    rc = SGPE_BY_DTYPE(p, run_row, p, p->state, p->state, true, true, dt_sub, true, 0, 0, 1.0, st);
    if (rc) return rc;
    p->phase = sgpe_plan::MID;
    p->pending_dt = dt_sub;
    p->scale_pending = false;
    return 0;
}

int close_junction(sgpe_plan* p, cudaStream_t st) {
    if (p->phase != sgpe_plan::MID) return 0;
    const bool want_e = p->pend_energy != nullptr && p->pend_eslot >= 0;
    int rc = SGPE_BY_DTYPE(p, run_col, p, p->state, p->state, true, true, p->pending_dt / 2, false, 0.0, false, 0, 0,
                           1.0, p->pend_pops, p->pend_stride, p->pend_slot, st, want_e ? p->scratch : nullptr);
    if (rc) return rc;
    if (want_e && (rc = energy_of_boundary(p, st))) return rc;
    p->pend_slot = -1;
    p->phase = sgpe_plan::KSPACE;
    p->scale_pending = true;
    return 0;
}

template <typename T>
int run_scale(sgpe_plan* p, const void* in, void* out, const double* totals, double atom_over_dv, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::ScaleArgs<T> a;
    a.in = static_cast<const C*>(in); a.out = static_cast<C*>(out);
    a.per_batch = 2 * p->plane; a.totals = totals; a.atom_over_dv = atom_over_dv;
    long long blocks = (a.per_batch + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dim3 grid((unsigned)blocks, p->batch), block(256);
    SGPE_LAUNCH((sgpe::scale_by_norm<T>), grid, block, 0, st, a);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int run_sumsq(sgpe_plan* p, const void* in, double* out2, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::SumsqArgs<T> a;
    a.in = static_cast<const C*>(in); a.plane = p->plane;
    a.partials = p->partials; a.counter = p->counter; a.totals = p->totals_aux; a.out2 = out2;
    long long blocks = (p->plane + 256 * 8 - 1) / (256 * 8);
    if (blocks > p->max_tiles) blocks = p->max_tiles;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, p->batch), block(256);
    SGPE_LAUNCH((sgpe::sumsq_pass<T>), grid, block, 32 * 2 * sizeof(double), st, a);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int run_kinetic(sgpe_plan* p, const void* psik, double* out, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::KineticArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.in = static_cast<const C*>(psik); a.nx = p->nx; a.ny = p->ny; a.plane = p->plane;
    a.kin_mode = p->kin_mode; a.kin0 = p->kin0; a.kin1 = p->kin1; a.kin_bstride = p->kin_bs;
    a.kin_x = p->kin_x; a.kin_y = p->kin_y; a.kinx_bstride = p->kin_xbs; a.kiny_bstride = p->kin_ybs;
    a.scale = p->dv_k;
    a.partials = p->partials; a.counter = p->counter; a.out = out;
    long long blocks = (p->plane + 256 * 8 - 1) / (256 * 8);
    if (blocks > p->max_tiles) blocks = p->max_tiles;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, p->batch), block(256);
    SGPE_LAUNCH((sgpe::kinetic_pass<T>), grid, block, 32 * 2 * sizeof(double), st, a);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int run_energy(sgpe_plan* p, const void* psi, int unwrap_mode, double kl_term, double* out, cudaStream_t st,
               long long out_bstride = 4, bool have_maxdens = false, bool polar = false, int* slot_ctr = nullptr) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::MaxDensArgs<T> m;
    m.psi = static_cast<const C*>(psi); m.plane = p->plane;
    m.partials = p->partials; m.counter = p->counter; m.maxdens = p->maxdens;
    long long blocks = (p->plane + 256 * 8 - 1) / (256 * 8);
    if (blocks > 1024) blocks = 1024;
    if (!have_maxdens) {
        dim3 grid((unsigned)blocks, p->batch), block(256);
        SGPE_LAUNCH((sgpe::maxdens_pass<T>), grid, block, 256 * 2 * sizeof(double), st, m);
        p->launches++;
        SGPE_CUDA(cudaGetLastError());
    }
    sgpe::EnergyArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.psi = static_cast<const C*>(psi); a.nx = p->nx; a.ny = p->ny; a.plane = p->plane;
    a.pot0 = p->pot0; a.pot1 = p->pot1; a.pot_bstride = p->pot_bs;
    a.pot_mode = p->pot_mode; a.pot_x = p->pot_x; a.pot_y = p->pot_y;
    a.potx_bstride = p->pot_xbs; a.poty_bstride = p->pot_ybs;
    a.cpl_mode = p->cpl_mode; a.coupling = p->cpl; a.cpl_bstride = p->cpl_bs; a.omega_b = p->omega;
    if (p->ecpl_mode >= 0) { a.cpl_mode = p->ecpl_mode; a.coupling = p->ecpl; a.cpl_bstride = p->ecpl_bs; a.omega_b = p->eomega; }
    a.g_uu = p->g_uu; a.g_dd = p->g_dd; a.g_ud = p->g_ud;
    a.kl2 = kl_term;
    a.inv_h0 = 1.0 / p->dx;        // np.gradient(f, dx, dy): dx goes with axis 0 (tensor_tools.py:342)
    a.inv_h1 = 1.0 / p->dy;
    a.unwrap_mode = unwrap_mode;
    a.inc = p->unwrap_inc;
    a.maxdens = p->maxdens; a.partials = p->partials; a.counter = p->counter; a.out = out;
    a.out_bstride = out_bstride;
    a.polar = polar ? 1 : 0;
    a.slot_ctr = slot_ctr;
    if (slot_ctr && p->energy_kernel == 1) return fail(SGPE_EINVAL, "the tiled energy kernel has no replay slot");
    if (polar && p->energy_kernel == 1) return fail(SGPE_EINVAL, "the tiled energy kernel takes (re, im) input");
    if (p->energy_kernel == 1) {       // the tiled kernel of round 1 (option "energy_kernel", cross-checks)
        long long tiles = (long long)((p->nx + 31) / 32) * ((p->ny + 7) / 8);
        blocks = tiles < 888 ? tiles : 888;                     // six 256-thread CTAs per SM x 148 SMs
        if (blocks > p->max_tiles / 2) blocks = p->max_tiles / 2;   // four partial sums per CTA in `partials`
        dim3 grid((unsigned)blocks, p->batch), block(256);
        SGPE_LAUNCH((sgpe::energy_pass<T>), grid, block, (32 * 4 + 2 * 34 * 10) * sizeof(double), st, a);
        p->launches++;
        SGPE_CUDA(cudaGetLastError());
        return 0;
    }
    if (polar && p->energy_polar == 1 && unwrap_mode != 2) {
        // warp-per-band stencil pass on (|psi|, arg psi): CTAs of 8 warps x 30 columns, bands of `rows` rows, three
        // resident CTAs per SM, at most max_tiles / 2 of them (four partial sums per CTA)
        const int ncb8 = ((p->nx + 29) / 30 + 7) / 8;
        long long want = 296 / ncb8;
        if (want > p->max_tiles / 2 / ncb8) want = p->max_tiles / 2 / ncb8;
        if (want < 1) want = 1;
        int rows = (int)((p->ny + want - 1) / want);
        if (rows < 8) rows = 8;
        if (rows > p->ny) rows = p->ny;
        a.rows = rows;
        const int nyb = (p->ny + rows - 1) / rows;
        dim3 grid((unsigned)(ncb8 * nyb), p->batch), block(256);
        if (unwrap_mode == 1) { SGPE_LAUNCH((sgpe::energy_polar_pass<T, true>), grid, block, 32 * 4 * sizeof(double), st, a); }
        else { SGPE_LAUNCH((sgpe::energy_polar_pass<T, false>), grid, block, 32 * 4 * sizeof(double), st, a); }
        p->launches++;
        SGPE_CUDA(cudaGetLastError());
        return 0;
    }
    // streaming kernel: bands of `rows` rows x 256 columns, ~4 CTAs per SM, at most max_tiles / 2 of them
    const int nxb = (p->nx + 255) / 256;
    long long want = 592 / nxb;
    if (want > p->max_tiles / 2 / nxb) want = p->max_tiles / 2 / nxb;
    if (want < 1) want = 1;
    int rows = (int)((p->ny + want - 1) / want);
    if (rows < 8) rows = 8;
    if (rows > p->ny) rows = p->ny;
    a.rows = rows;
    const int nyb = (p->ny + rows - 1) / rows;
    dim3 grid((unsigned)(nxb * nyb), p->batch), block(256);
    if (polar) { SGPE_LAUNCH((sgpe::energy_stream_pass<T, true>), grid, block, (32 * 4 + 2 * 4 * 258) * sizeof(double), st, a); }
    else { SGPE_LAUNCH((sgpe::energy_stream_pass<T, false>), grid, block, (32 * 4 + 2 * 4 * 258) * sizeof(double), st, a); }
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

// ---- two-dimensional phase unwrapping (Herraez et al. 2002; skimage.restoration.unwrap_phase at the reference's
// tensor_tools.py:531).  Device: angle, reliabilities, edge keys, radix sort, the region merging (spanning tree and
// surviving group, unwrap.cuh) and the final pass, steered from unwrap_increments_impl below.  The sequential
// edge-by-edge merging that defines the result is kept as the cross-check (option "unwrap_merge" = 1):
//
// Groups are an offset-carrying union-find: pot[x] = increment(x) - increment(parent[x]); a root keeps the
// increments of its group fixed (its own is 0) and the absorbed root receives the shift of its whole group, which is
// what the published algorithm does by walking the smaller group's pixel list.  Who absorbs whom follows the
// published rules (a lone second pixel joins the first pixel's group, a lone first pixel the second's, otherwise the
// strictly larger group absorbs and a tie goes to the second pixel's group), so the integer field — global offset
// included — equals that of the list-based formulation (oracle/unwrap_herraez.c) for the same edge order.
struct UnwrapForest {
    struct Node { int32_t parent, pot; };          // one cache line access per hop
    std::vector<Node> node;
    std::vector<int32_t> size;
    explicit UnwrapForest(size_t n) : node(n), size(n, 1) {
        for (size_t i = 0; i < n; i++) node[i] = {(int32_t)i, 0};
    }
    int32_t find(int32_t x, int32_t* offset) {
        int32_t r = x, sum = 0;
        while (node[r].parent != r) { sum += node[r].pot; r = node[r].parent; }
        int32_t cur = x, s = sum;
        while (cur != r && node[cur].parent != r) {   // path compression, offsets re-expressed against the root
            const int32_t next = node[cur].parent, own = node[cur].pot;
            node[cur] = {r, s};
            s -= own; cur = next;
        }
        *offset = sum;
        return r;
    }
};

void unwrap_merge(int nx, int ny, const uint32_t* order, size_t n_edges, int32_t* inc) {
    const size_t plane = (size_t)nx * ny;
    const size_t n_horizontal = (size_t)ny * (nx - 1);
    UnwrapForest f(plane);
    for (size_t k = 0; k < n_edges; k++) {
        const uint32_t e = order[k] >> 2;
        const int32_t wraps = (int32_t)(order[k] & 3u) - 1;
        int32_t p1, p2;
        if (e < n_horizontal) { const uint32_t i = e / (uint32_t)(nx - 1), j = e - i * (uint32_t)(nx - 1); p1 = (int32_t)(i * (uint32_t)nx + j); p2 = p1 + 1; }
        else { p1 = (int32_t)(e - n_horizontal); p2 = p1 + nx; }
        int32_t a1, a2;
        const int32_t r1 = f.find(p1, &a1), r2 = f.find(p2, &a2);
        if (r1 == r2) continue;
        const int32_t s1 = f.size[r1], s2 = f.size[r2];
        const bool second_joins = (s2 == 1) || (s1 != 1 && s1 > s2);
        if (second_joins) { f.node[r2] = {r1, a1 - wraps - a2}; f.size[r1] = s1 + s2; }
        else              { f.node[r1] = {r2, a2 + wraps - a1}; f.size[r2] = s1 + s2; }
    }
    for (size_t i = 0; i < plane; i++) { int32_t a; f.find((int32_t)i, &a); inc[i] = a; }
}

unsigned unwrap_blocks(long long n) { long long b = (n + 255) / 256; return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b)); }

int unwrap_increments_impl(sgpe_plan* p, const double* phi_dev, int nplanes, int* inc_dev, cudaStream_t st);

// phi_dev: nplanes x [ny][nx] wrapped phases -> inc_dev: nplanes x [ny][nx] multiples of 2 pi.  Synchronises st.
// (host containers and threads live behind this call: nothing may throw across the C boundary)
int unwrap_increments(sgpe_plan* p, const double* phi_dev, int nplanes, int* inc_dev, cudaStream_t st) {
    try {
        return unwrap_increments_impl(p, phi_dev, nplanes, inc_dev, st);
    } catch (const std::bad_alloc&) {
        return fail(SGPE_ENOMEM, "phase unwrapping: host allocation failed");
    } catch (const std::exception& e) {
        return fail(SGPE_ECUDA, std::string("phase unwrapping: ") + e.what());
    }
}

// host staging of the phase unwrapping: plain pageable memory.  (Page-locked buffers were measured: cudaHostAlloc of the
// ~100 MB costs 50 ms per plan, more than the pageable copies of 4 ms worth of device work ever did; the evaluation is
// bound by the host-side region merging, ~245 ms per 2048^2 plane on the GPU box's cores.)
static void* host_pinned(size_t bytes) { return malloc(bytes); }
static void host_pinned_free(void* q) { free(q); }

void unwrap_release(sgpe_plan::UnwrapCache& w) {
    cudaFree(w.rel); cudaFree(w.keys); cudaFree(w.keys_sorted); cudaFree(w.vals); cudaFree(w.vals_sorted); cudaFree(w.tmp);
    cudaFree(w.counters); cudaFree(w.ctrl);
    host_pinned_free(w.order_host); host_pinned_free(w.inc_host);
    w = sgpe_plan::UnwrapCache();
}

#ifndef SGPE_EMU
struct UnwrapIsTreeEdge { __device__ bool operator()(unsigned x) const { return x != sgpe::kUnwrapNoEdge; } };
#endif

// the scratch of the phase unwrapping lives with the plan (device buffers, the sort's and the selection's work space,
// host staging for the edge order going down and the increments coming up): nothing is allocated per evaluation.
// The device-side merging reuses the sort's buffers once the order is known: keys -> forest nodes, then the tree-edge
// list; keys_sorted -> pending hooks, then the flagged candidates; vals -> rank of every edge; rel -> best edge per group.
static int unwrap_scratch(sgpe_plan* p, int nplanes, size_t plane, size_t n_edges, bool device_sort, cudaStream_t st) {
    sgpe_plan::UnwrapCache& w = p->unwrap;
    if (w.nplanes >= nplanes && w.plane == plane && w.device_sort == device_sort && w.rel) return 0;
    unwrap_release(w);
    if (cudaMalloc((void**)&w.rel, plane * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&w.keys, n_edges * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc((void**)&w.vals, n_edges * sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc((void**)&w.keys_sorted, n_edges * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc((void**)&w.vals_sorted, n_edges * sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc((void**)&w.counters, 16 * sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc((void**)&w.ctrl, 4 * sizeof(long long)) != cudaSuccess)
        return fail(SGPE_ENOMEM, "unwrap scratch allocation failed");
#ifndef SGPE_EMU
    size_t select_bytes = 0;
    SGPE_CUDA(cub::DeviceSelect::If(nullptr, select_bytes, w.vals, w.vals_sorted, w.counters, (long long)n_edges,
                                    UnwrapIsTreeEdge(), st));
    w.tmp_bytes = 0;
    if (device_sort)
        SGPE_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, w.tmp_bytes, w.keys, w.keys_sorted, w.vals, w.vals_sorted,
                                                  (long long)n_edges, 0, 64, st));
    w.tmp_bytes = std::max(w.tmp_bytes, select_bytes);
    if (cudaMalloc(&w.tmp, w.tmp_bytes ? w.tmp_bytes : 1) != cudaSuccess)
        return fail(SGPE_ENOMEM, "unwrap scratch allocation failed");
#else
    (void)st;
#endif
    w.order_host = static_cast<uint32_t*>(host_pinned((size_t)nplanes * n_edges * sizeof(uint32_t)));
    w.inc_host = static_cast<int32_t*>(host_pinned((size_t)nplanes * plane * sizeof(int32_t)));
    if (!w.order_host || !w.inc_host) return fail(SGPE_ENOMEM, "unwrap host staging allocation failed");
    w.nplanes = nplanes; w.plane = plane; w.device_sort = device_sort;
    return 0;
}

// The pixel whose group never moved: Kruskal over the spanning tree's edges in rank order with the published rules for
// who absorbs whom (see UnwrapForest above), sizes only — the absorbed root hangs below the surviving one, so the
// root of the final tree IS a pixel of the group that kept its values throughout.  tree[k] = first pixel of the k-th
// tree edge, bit 31 set for a vertical edge (unwrap_tree_edges_pass).
int64_t unwrap_anchor(int nx, const uint32_t* tree, size_t n_tree, size_t plane) {
    struct Node { int32_t parent, size; };
    std::vector<Node> nd(plane);
    for (size_t i = 0; i < plane; i++) nd[i] = {(int32_t)i, 1};
    Node* n = nd.data();
    auto find = [n](int32_t x) {
        while (n[x].parent != x) { const int32_t g = n[n[x].parent].parent; n[x].parent = g; x = g; }   // path halving
        return x;
    };
    constexpr size_t kAhead = 24;
    for (size_t k = 0; k < n_tree; k++) {
        if (k + kAhead < n_tree) {
            const uint32_t t = tree[k + kAhead];
            const uint32_t q = t & 0x7fffffffu;
            __builtin_prefetch(&n[q]);
            if (t >> 31) __builtin_prefetch(&n[q + (uint32_t)nx]);
        }
        const uint32_t t = tree[k];
        const int32_t p1 = (int32_t)(t & 0x7fffffffu), p2 = p1 + ((t >> 31) ? nx : 1);
        const int32_t r1 = find(p1), r2 = find(p2);
        if (r1 == r2) continue;                                  // (cannot happen for tree edges)
        const int32_t s1 = n[r1].size, s2 = n[r2].size;
        const bool second_joins = (s2 == 1) || (s1 != 1 && s1 > s2);
        if (second_joins) { n[r2].parent = r1; n[r1].size = s1 + s2; }
        else              { n[r1].parent = r2; n[r2].size = s1 + s2; }
    }
    return find(0);
}

// The same pixel found on the device (unwrap.cuh, "the pixel group that never moves"): per level a bisection over the
// rank threshold with a lock-free union-find, until the group is small enough for the host pass above.  The tree-edge
// list is in w.keys (as left by the selection); forests, pixel lists and edge lists live in the sort's buffers.
static int unwrap_anchor_device(sgpe_plan* p, int nx, size_t plane, cudaStream_t st, int64_t* anchor, int* levels, int* probes) {
    sgpe_plan::UnwrapCache& w = p->unwrap;
    unsigned* tree = reinterpret_cast<unsigned*>(w.keys);
    unsigned* size_a = tree + plane;                 // group sizes at the roots of forest a / b
    unsigned* size_b = size_a + plane;
    unsigned* forest_a = reinterpret_cast<unsigned*>(w.rel);
    unsigned* forest_b = forest_a + plane;
    unsigned* vl[2] = {w.vals, w.vals + plane};
    sgpe::UnwrapEdge* el[2] = {reinterpret_cast<sgpe::UnwrapEdge*>(w.keys_sorted), reinterpret_cast<sgpe::UnwrapEdge*>(w.keys_sorted) + plane};
    unsigned* result = w.counters + 8;
    long long* ctrl = w.ctrl;
    struct { long long ctrl[4]; unsigned result[8]; } back;
    const dim3 block(256);
    long long nv = (long long)plane, ne = nv - 1, bound = ne;
    int cur = 0;
    SGPE_LAUNCH((sgpe::unwrap_level0_pass), dim3(unwrap_blocks(nv)), block, 0, st, tree, nv, vl[0], el[0]);
    p->launches++;
    while (ne > (long long)p->unwrap_tail) {
        const dim3 grid_v(unwrap_blocks(nv)), grid_e(unwrap_blocks(ne));
        // no group of more than nv / 2 pixels with the edges <= lo (none at all: lo = -1), one with those <= hi
        SGPE_LAUNCH((sgpe::unwrap_level_begin_pass), dim3(1), dim3(32), 0, st, -1ll, bound - 1, ctrl, result);
        SGPE_LAUNCH((sgpe::unwrap_level_reset_pass), grid_v, block, 0, st, vl[cur], nv, (const long long*)nullptr, forest_a, forest_b, size_a, size_b);
        p->launches += 2;
        int enqueued = 0;
        for (long long len = bound; len > 1; len = (len + 1) / 2) {        // the interval shrinks to ceil(len / 2) at worst
            SGPE_LAUNCH((sgpe::unwrap_level_reset_pass), grid_v, block, 0, st, vl[cur], nv, (const long long*)ctrl, forest_a, forest_b, size_a, size_b);
            SGPE_LAUNCH((sgpe::unwrap_level_union_pass), grid_e, block, 0, st, el[cur], ne, (const long long*)ctrl, nx, forest_a, forest_b, size_a, size_b, nv, result);
            SGPE_LAUNCH((sgpe::unwrap_level_step_pass), dim3(1), dim3(32), 0, st, nv, ctrl, result);
            p->launches += 3;
            enqueued++;
        }
        *probes += enqueued;
        // edge `hi` creates the majority group; the forest at lo = hi - 1 is the state just before it
        SGPE_LAUNCH((sgpe::unwrap_level_sides_pass), grid_e, block, 0, st, el[cur], ne, (const long long*)ctrl, nx, forest_a, forest_b, size_a, size_b, result);
        p->launches++;
        SGPE_CUDA(cudaGetLastError());
        SGPE_CUDA(cudaMemcpyAsync(back.ctrl, ctrl, sizeof(back.ctrl), cudaMemcpyDeviceToHost, st));
        SGPE_CUDA(cudaMemcpyAsync(back.result, result, sizeof(back.result), cudaMemcpyDeviceToHost, st));
        SGPE_CUDA(cudaStreamSynchronize(st));
        const long long hi = back.ctrl[1];
        const long long s1 = back.result[2], s2 = back.result[4];
        if (back.ctrl[1] - back.ctrl[0] != 1 || s1 < 1 || s2 < 1 || 2 * (s1 + s2) <= nv || 2 * s1 > nv || 2 * s2 > nv)
            return fail(SGPE_ECUDA, "phase unwrapping: anchor bisection lost its edge (internal error)");
        const bool second_joins = (s2 == 1) || (s1 != 1 && s1 > s2);
        const unsigned keep = second_joins ? back.result[1] : back.result[3];
        const long long keep_size = second_joins ? s1 : s2;
        SGPE_LAUNCH((sgpe::unwrap_level_select_pass), dim3(unwrap_blocks(nv)), block, 0, st, vl[cur], nv, el[cur], ne, (unsigned)hi, keep,
                    back.ctrl[2] ? forest_b : forest_a, vl[1 - cur], el[1 - cur], result);
        p->launches++;
        SGPE_CUDA(cudaMemcpyAsync(back.result, result, sizeof(back.result), cudaMemcpyDeviceToHost, st));
        SGPE_CUDA(cudaStreamSynchronize(st));
        if ((long long)back.result[5] != keep_size || (long long)back.result[6] != keep_size - 1)
            return fail(SGPE_ECUDA, "phase unwrapping: anchor bisection selected a wrong group (internal error)");
        nv = keep_size; ne = nv - 1; bound = hi; cur ^= 1;
        (*levels)++;
    }
    if (ne == 0) {
        unsigned v = 0;
        SGPE_CUDA(cudaMemcpyAsync(&v, vl[cur], sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        SGPE_CUDA(cudaStreamSynchronize(st));
        *anchor = (int64_t)v;
        return 0;
    }
    // the rest on the host: the group's tree edges in rank order, pixels renumbered densely
    std::vector<sgpe::UnwrapEdge> edges((size_t)ne);
    SGPE_CUDA(cudaMemcpyAsync(edges.data(), el[cur], (size_t)ne * sizeof(sgpe::UnwrapEdge), cudaMemcpyDeviceToHost, st));
    SGPE_CUDA(cudaStreamSynchronize(st));
    std::sort(edges.begin(), edges.end(), [](const sgpe::UnwrapEdge& a, const sgpe::UnwrapEdge& b) { return a.k < b.k; });
    std::vector<uint32_t> pixels;
    pixels.reserve(2 * (size_t)ne);
    for (const auto& e : edges) {
        const uint32_t p1 = e.t & 0x7fffffffu;
        pixels.push_back(p1);
        pixels.push_back(p1 + ((e.t >> 31) ? (uint32_t)nx : 1u));
    }
    std::sort(pixels.begin(), pixels.end());
    pixels.erase(std::unique(pixels.begin(), pixels.end()), pixels.end());
    auto dense = [&pixels](uint32_t v) { return (int32_t)(std::lower_bound(pixels.begin(), pixels.end(), v) - pixels.begin()); };
    struct Node { int32_t parent, size; };
    std::vector<Node> nd(pixels.size());
    for (size_t i = 0; i < nd.size(); i++) nd[i] = {(int32_t)i, 1};
    auto find = [&nd](int32_t x) { while (nd[x].parent != x) { const int32_t g = nd[nd[x].parent].parent; nd[x].parent = g; x = g; } return x; };
    for (const auto& e : edges) {
        const uint32_t p1 = e.t & 0x7fffffffu;
        const int32_t r1 = find(dense(p1)), r2 = find(dense(p1 + ((e.t >> 31) ? (uint32_t)nx : 1u)));
        if (r1 == r2) continue;
        const int32_t s1 = nd[r1].size, s2 = nd[r2].size;
        const bool second_joins = (s2 == 1) || (s1 != 1 && s1 > s2);
        if (second_joins) { nd[r2].parent = r1; nd[r1].size = s1 + s2; }
        else              { nd[r1].parent = r2; nd[r2].size = s1 + s2; }
    }
    *anchor = (int64_t)pixels[(size_t)find(0)];
    return 0;
}

int unwrap_increments_impl(sgpe_plan* p, const double* phi_dev, int nplanes, int* inc_dev, cudaStream_t st) {
    const int nx = p->nx, ny = p->ny;
    const size_t plane = (size_t)p->plane;
    const size_t n_edges = (size_t)ny * (nx - 1) + (size_t)nx * (ny - 1);
    if (n_edges >= (1ull << 29)) return fail(SGPE_EINVAL, "phase unwrapping: mesh too large (edge ids are 29 bits)");
    bool device_sort = false;
#ifndef SGPE_EMU
    device_sort = p->unwrap_sort == 0;
#endif
    const bool device_merge = p->unwrap_merge == 0 && plane > 1;
    // the anchor search on the device handles the planes one after the other; the host pass runs one plane per core
    const bool device_anchor = device_merge && n_edges >= plane + plane / 2 + 2 &&       // (room for the lists in the sort's buffers)
                               (p->unwrap_anchor == 1 || (p->unwrap_anchor < 0 && nplanes <= 4 && plane >= 65536));
    const bool timing = getenv("SGPE_UNWRAP_TIMING") != nullptr;       // dev: phase times on stderr
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    int rc = unwrap_scratch(p, nplanes, plane, n_edges, device_sort, st);
    if (rc) return rc;
    if (timing) { cudaStreamSynchronize(st); fprintf(stderr, "unwrap: scratch %.1f ms\n", now() - t_begin); }
    sgpe_plan::UnwrapCache& w = p->unwrap;
    std::vector<unsigned long long> host_keys;
    std::vector<int64_t> anchors((size_t)nplanes, 0);
    // The sequential part of a plane (host) starts as soon as its edges have arrived and runs beside the device work of
    // the next plane; the planes are independent.
    std::vector<std::thread> pool;
    struct Joiner { std::vector<std::thread>& t; ~Joiner() { for (auto& th : t) if (th.joinable()) th.join(); } } joiner{pool};
    const unsigned max_threads = std::max(1u, std::thread::hardware_concurrency());
    const dim3 grid_plane(unwrap_blocks((long long)plane)), grid_edges(unwrap_blocks((long long)n_edges)), block(256);
    for (int pl = 0; pl < nplanes; pl++) {
        const double* phi = phi_dev + (size_t)pl * plane;
        uint32_t* order = w.order_host + (size_t)pl * n_edges;
        SGPE_LAUNCH((sgpe::unwrap_reliab_pass), grid_plane, block, 0, st, phi, nx, ny, w.rel);
        SGPE_LAUNCH((sgpe::unwrap_edge_pass), grid_plane, block, 0, st, phi, w.rel, nx, ny, w.keys, w.vals);
        p->launches += 2;
        SGPE_CUDA(cudaGetLastError());
        if (device_sort) {
#ifndef SGPE_EMU
            // least-significant-digit radix sort: stable, so equal keys stay in edge-id order
            SGPE_CUDA(cub::DeviceRadixSort::SortPairs(w.tmp, w.tmp_bytes, w.keys, w.keys_sorted, w.vals, w.vals_sorted,
                                                      (long long)n_edges, 0, 64, st));
            p->launches++;
            if (!device_merge) {
                SGPE_CUDA(cudaMemcpyAsync(order, w.vals_sorted, n_edges * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
                SGPE_CUDA(cudaStreamSynchronize(st));
            }
#endif
        } else {
            host_keys.resize(n_edges);
            std::vector<uint32_t> vals(n_edges);
            SGPE_CUDA(cudaMemcpyAsync(host_keys.data(), w.keys, n_edges * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            SGPE_CUDA(cudaMemcpyAsync(vals.data(), w.vals, n_edges * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            SGPE_CUDA(cudaStreamSynchronize(st));
            std::vector<uint32_t> perm(n_edges);
            for (size_t k = 0; k < n_edges; k++) perm[k] = (uint32_t)k;
            const unsigned long long* hk = host_keys.data();
            std::sort(perm.begin(), perm.end(), [hk](uint32_t a, uint32_t b) { return hk[a] != hk[b] ? hk[a] < hk[b] : a < b; });
            for (size_t k = 0; k < n_edges; k++) order[k] = vals[perm[k]];
            if (device_merge)
                SGPE_CUDA(cudaMemcpyAsync(w.vals_sorted, order, n_edges * sizeof(unsigned), cudaMemcpyHostToDevice, st));
        }
        if (timing) { cudaStreamSynchronize(st); fprintf(stderr, "unwrap: plane %d order ready at %.1f ms\n", pl, now() - t_begin); }
        int32_t* inc = w.inc_host + (size_t)pl * plane;
        if (!device_merge) {
            if (max_threads > 1 && nplanes > 1) {
                if (pool.size() >= max_threads) { pool.front().join(); pool.erase(pool.begin()); }
                pool.emplace_back([=]() {
                    const double t0 = now();
                    unwrap_merge(nx, ny, order, n_edges, inc);
                    if (timing) fprintf(stderr, "unwrap: merge of plane %d %.1f ms\n", pl, now() - t0);
                });
            } else {
                unwrap_merge(nx, ny, order, n_edges, inc);
            }
            continue;
        }
        // ---- spanning tree and relative increments on the device (unwrap.cuh)
        unsigned long long* node = w.keys;
        unsigned long long* pend = w.keys_sorted;
        unsigned* rank_of = w.vals;
        unsigned* best = reinterpret_cast<unsigned*>(w.rel);
        int* raw = inc_dev + (size_t)pl * plane;
        SGPE_LAUNCH((sgpe::unwrap_rank_pass), grid_edges, block, 0, st, w.vals_sorted, (long long)n_edges, rank_of);
        SGPE_LAUNCH((sgpe::unwrap_forest_init_pass), grid_plane, block, 0, st, (long long)plane, node, best);
        p->launches += 2;
        size_t joined = 0;
        int rounds = 0;
        while (joined + 1 < plane) {
            unsigned hooked = 0;
            SGPE_CUDA(cudaMemsetAsync(w.counters, 0, sizeof(unsigned), st));
            SGPE_LAUNCH((sgpe::unwrap_minedge_pass), grid_plane, block, 0, st, node, rank_of, nx, ny, best);
            SGPE_LAUNCH((sgpe::unwrap_hook_pass), grid_plane, block, 0, st, node, best, w.vals_sorted, nx, ny, pend, w.counters);
            SGPE_LAUNCH((sgpe::unwrap_adopt_pass), grid_plane, block, 0, st, (long long)plane, pend, node, best);
            SGPE_LAUNCH((sgpe::unwrap_compress_pass), grid_plane, block, 0, st, (long long)plane, node);
            p->launches += 4;
            SGPE_CUDA(cudaMemcpyAsync(&hooked, w.counters, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            SGPE_CUDA(cudaStreamSynchronize(st));
            rounds++;
            if (hooked == 0) return fail(SGPE_ECUDA, "phase unwrapping: the spanning tree did not close (internal error)");
            joined += hooked;
        }
        if (joined + 1 != plane) return fail(SGPE_ECUDA, "phase unwrapping: wrong number of tree edges (internal error)");
        SGPE_LAUNCH((sgpe::unwrap_offsets_pass), grid_plane, block, 0, st, (long long)plane, node, raw);
        unsigned* cand = reinterpret_cast<unsigned*>(w.keys_sorted);
        unsigned* tree_dev = reinterpret_cast<unsigned*>(w.keys);
        SGPE_LAUNCH((sgpe::unwrap_tree_edges_pass), grid_edges, block, 0, st, w.vals_sorted, (long long)n_edges, nx, ny, cand);
        p->launches += 2;
        SGPE_CUDA(cudaGetLastError());
#ifndef SGPE_EMU
        SGPE_CUDA(cub::DeviceSelect::If(w.tmp, w.tmp_bytes, cand, tree_dev, w.counters + 1, (long long)n_edges,
                                        UnwrapIsTreeEdge(), st));
        p->launches++;
#else
        { size_t m = 0; for (size_t k = 0; k < n_edges; k++) if (cand[k] != sgpe::kUnwrapNoEdge) tree_dev[m++] = cand[k]; w.counters[1] = (unsigned)m; }
#endif
        unsigned n_tree = 0;
        SGPE_CUDA(cudaMemcpyAsync(&n_tree, w.counters + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        if (!device_anchor)
            SGPE_CUDA(cudaMemcpyAsync(order, tree_dev, (plane - 1) * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        SGPE_CUDA(cudaStreamSynchronize(st));
        if ((size_t)n_tree + 1 != plane) return fail(SGPE_ECUDA, "phase unwrapping: tree edge list incomplete (internal error)");
        if (timing) fprintf(stderr, "unwrap: plane %d tree ready at %.1f ms (%d rounds)\n", pl, now() - t_begin, rounds);
        if (device_anchor) {
            int levels = 0, probes = 0;
            rc = unwrap_anchor_device(p, nx, plane, st, &anchors[(size_t)pl], &levels, &probes);
            if (rc) return rc;
            if (timing) fprintf(stderr, "unwrap: plane %d anchor %lld at %.1f ms (%d levels, %d probes)\n", pl,
                                (long long)anchors[(size_t)pl], now() - t_begin, levels, probes);
            continue;
        }
        int64_t* anchor = &anchors[(size_t)pl];
        auto work = [=]() {
            const double t0 = now();
            *anchor = unwrap_anchor(nx, order, plane - 1, plane);
            if (timing) fprintf(stderr, "unwrap: anchor pass of plane %d %.1f ms\n", pl, now() - t0);
        };
        if (max_threads > 1 && nplanes > 1) {
            if (pool.size() >= max_threads) { pool.front().join(); pool.erase(pool.begin()); }
            pool.emplace_back(work);
        } else {
            work();
        }
    }
    for (auto& th : pool) th.join();
    pool.clear();
    if (timing) fprintf(stderr, "unwrap: merged at %.1f ms\n", now() - t_begin);
    if (!device_merge) {
        SGPE_CUDA(cudaMemcpyAsync(inc_dev, w.inc_host, (size_t)nplanes * plane * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    } else {
        for (int pl = 0; pl < nplanes; pl++) {
            int* raw = inc_dev + (size_t)pl * plane;
            SGPE_LAUNCH((sgpe::unwrap_anchor_pass), grid_plane, block, 0, st, (long long)plane, (long long)anchors[(size_t)pl], raw);
            SGPE_LAUNCH((sgpe::unwrap_anchor_zero_pass), dim3(1), dim3(32), 0, st, (long long)anchors[(size_t)pl], raw);
            p->launches += 2;
        }
        SGPE_CUDA(cudaGetLastError());
    }
    SGPE_CUDA(cudaStreamSynchronize(st));
    return 0;
}

template <typename T>
int run_unwrap_angles(sgpe_plan* p, const void* psi, long long total, double* phi, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    SGPE_LAUNCH((sgpe::unwrap_angle_pass<T>), dim3(unwrap_blocks(total)), dim3(256), 0, st, static_cast<const C*>(psi), total, phi);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int run_unwrap_apply(sgpe_plan* p, const void* psi, const double* phi, const int* inc, int nplanes, bool mask,
                     unsigned long long* maxbits, double* out, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    const long long total = (long long)nplanes * p->plane;
    const C* z = mask ? static_cast<const C*>(psi) : nullptr;
    if (mask) {
        SGPE_CUDA(cudaMemsetAsync(maxbits, 0, sizeof(unsigned long long) * nplanes, st));
        unsigned blocks = unwrap_blocks(p->plane);
        if (blocks > 256) blocks = 256;
        SGPE_LAUNCH((sgpe::unwrap_maxdens_pass<T>), dim3(blocks, (unsigned)nplanes), dim3(256), 256 * sizeof(double), st, z,
                    p->plane, maxbits);
        p->launches++;
    }
    SGPE_LAUNCH((sgpe::unwrap_apply_pass<T>), dim3(unwrap_blocks(total)), dim3(256), 0, st, z, phi, inc, maxbits, p->plane,
                total, out);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int run_klines(sgpe_plan* p, void* buf, bool fwd, bool has_a, double tau_a, bool has_b, double tau_b, bool inv,
                      double* sums, bool scatter, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::KLineArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.in = static_cast<const C*>(buf); a.out = static_cast<C*>(buf);
    a.tw = static_cast<const C*>(p->tw_x);
    const int len = (p->n1 > 1) ? p->n2 : p->nx;           // contiguous transform length (sub-lines of a long line)
    a.nx = len; a.ny = p->ny * p->n1; a.plane = p->plane; a.group = p->n1;
    a.do_fwd = fwd; a.do_inv = inv; a.has_a = has_a; a.has_b = has_b;
    a.kin_mode = p->kin_mode;
    if (p->n1 > 1 && p->kin_mode == 0 && (has_a || has_b))
        return fail(SGPE_EINVAL, "long lines need the separable kinetic operator");
    if (has_a || has_b) {
        if (p->kin_mode == 0) {
            a.kin0 = p->kin0; a.kin1 = p->kin1;
            time_arg(p->tm, tau_a, &a.ka_re, &a.ka_im);
            time_arg(p->tm, tau_b, &a.kb_re, &a.kb_im);
        } else {
            sgpe_plan::FactorTable* t = nullptr;
            int rc;
            if (has_a) {
                if ((rc = factor_table<T>(p, p->kin_tab, 6, p->kin_x, 0, p->kin_y, 0, tau_a, st, &t))) return rc;
                a.pa = static_cast<const C*>(t->x); a.la = static_cast<const C*>(t->y);
            }
            if (has_b) {
                if ((rc = factor_table<T>(p, p->kin_tab, 6, p->kin_x, 0, p->kin_y, 0, tau_b, st, &t))) return rc;
                a.pb = static_cast<const C*>(t->x); a.lb = static_cast<const C*>(t->y);
            }
        }
    }
    const int wfirst = p->win.count ? p->win.first : 0, wcount = p->win.count ? p->win.count : p->ny;
    a.line0 = wfirst * p->n1; a.wlines = wcount * p->n1; a.max_ctas = p->win.max_ctas;
    a.partials = p->partials + 4LL * p->win.chunk * a.wlines; a.counter = p->counter + p->win.chunk; a.sums = sums;
    fill_scatter(p, scatter, &a.sc);
    ProfScope prof(p, 0, st);
    int rc = sgpe::launch_kline(len, p->dtype, p->tm, &a, st);
    if (rc != 0) return fail(SGPE_EINVAL, "line pass: unsupported geometry");
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int run_mid(sgpe_plan* p, void* buf, bool pre_tw, bool inv, bool pw, double dt_sub, bool fwd, bool post_tw,
                   const double* totals, double global_points, int inner, bool scatter, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::MidArgs<T> m;
    memset(&m, 0, sizeof(m));
    sgpe::RowArgs<T>& a = m.r;
    a.in = static_cast<const C*>(buf); a.out = static_cast<C*>(buf);
    a.tw = static_cast<const C*>(p->tw_mid);
    a.nx = p->nx; a.ny = p->ny; a.plane = p->plane;
    a.do_inv = inv; a.do_pw = pw; a.do_fwd = fwd; a.scale_out = 1.0;
    a.pot0 = p->pot0; a.pot1 = p->pot1; a.pot_bstride = 0;
    a.pot_mode = pw ? p->pot_mode : 0;
    if (pw && p->pot_mode == 1) {
        sgpe_plan::FactorTable* t = nullptr;
        int rc0 = factor_table<T>(p, p->pot_tab, 4, p->pot_x, 0, p->pot_y, 0, dt_sub, st, &t);
        if (rc0) return rc0;
        a.px = static_cast<const C*>(t->x); a.py = static_cast<const C*>(t->y);
    }
    a.cpl_mode = p->cpl_mode; a.coupling = p->cpl; a.omega_b = p->omega;
    a.eiphi = static_cast<const C*>(p->eiphi);
    a.g_uu = p->g_uu; a.g_dd = p->g_dd; a.g_ud = p->g_ud;
    time_arg(p->tm, dt_sub / 2, &a.ti_re, &a.ti_im);
    time_arg(p->tm, dt_sub, &a.tp_re, &a.tp_im);
    a.tc = dt_sub / 4;
    a.totals = totals;
    a.norm_c = p->atom_num / (p->dv_r * global_points);
    m.n2 = p->n2; m.pre_tw = pre_tw; m.post_tw = post_tw; m.tw4 = static_cast<const C*>(p->tw4);
    m.inner = 1;
    m.y0 = m.x0 = 0; m.wcount = p->ny; m.max_ctas = p->win.max_ctas;
    if (p->win.count) {
        m.wcount = p->win.count;
        if (inner > 1) m.x0 = p->win.first; else m.y0 = p->win.first;
    }
    if (inner > 1) {      // row-major k slab [2][n1][n2][inner]: one "line" whose contiguous dimension is n2 * inner
        if (pw) return fail(SGPE_EINVAL, "the point-wise operators are not available on the column-slab layout");
        if (inner != p->ny) return fail(SGPE_EINVAL, "inner must equal the number of lines of the plan");
        m.inner = inner; m.n2 = p->n2 * inner;
        a.ny = 1; a.nx = p->nx * inner;
    }
    fill_scatter(p, scatter, &a.sc);
    ProfScope prof(p, 1, st);
    int rc = sgpe::launch_mid(p->n1, p->dtype, p->tm, &m, st);
    if (rc != 0) return fail(SGPE_EINVAL, "strided pass: unsupported geometry");
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

// k-space junction on the row-major k slab [2][len][nlines] of a line plan (len = p->nx points down the columns,
// nlines = p->ny adjacent columns): the fused-exchange counterpart of run_klines.
template <typename T>
static int run_kcols(sgpe_plan* p, void* buf, bool fwd, bool has_a, double tau_a, bool has_b, double tau_b, bool inv,
                     double* sums, bool scatter, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::KColArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.in = static_cast<const C*>(buf); a.out = static_cast<C*>(buf);
    a.tw = static_cast<const C*>(p->tw_x);
    const int len = (p->n1 > 1) ? p->n2 : p->nx;
    a.inner = p->ny; a.groups = p->n1; a.plane = p->plane;
    a.do_fwd = fwd; a.do_inv = inv; a.has_a = has_a; a.has_b = has_b;
    a.kin_mode = p->kin_mode;
    if (scatter && p->n1 > 1) return fail(SGPE_EINVAL, "with a four-step split the strided pass stores the exchange");
    if (p->n1 > 1 && p->kin_mode == 0 && (has_a || has_b))
        return fail(SGPE_EINVAL, "long lines need the separable kinetic operator");
    if (has_a || has_b) {
        if (p->kin_mode == 0) {
            a.kin0 = p->kin0; a.kin1 = p->kin1;
            time_arg(p->tm, tau_a, &a.ka_re, &a.ka_im);
            time_arg(p->tm, tau_b, &a.kb_re, &a.kb_im);
        } else {
            sgpe_plan::FactorTable* t = nullptr;
            int rc;
            if (has_a) {
                if ((rc = factor_table<T>(p, p->kin_tab, 6, p->kin_x, 0, p->kin_y, 0, tau_a, st, &t))) return rc;
                a.pa = static_cast<const C*>(t->x); a.la = static_cast<const C*>(t->y);
            }
            if (has_b) {
                if ((rc = factor_table<T>(p, p->kin_tab, 6, p->kin_x, 0, p->kin_y, 0, tau_b, st, &t))) return rc;
                a.pb = static_cast<const C*>(t->x); a.lb = static_cast<const C*>(t->y);
            }
        }
    }
    a.x0 = p->win.count ? p->win.first : 0; a.wcount = p->win.count ? p->win.count : p->ny;
    a.partials = p->partials + 4LL * p->win.chunk * a.wcount * a.groups; a.counter = p->counter + p->win.chunk;
    a.sums = sums;
    fill_scatter(p, scatter, &a.sc);
    ProfScope prof(p, 0, st);
    int rc = sgpe::launch_kcol(len, p->dtype, p->tm, &a, st);
    if (rc != 0) return fail(SGPE_EINVAL, "column-slab pass: unsupported geometry");
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int upload_tw4(sgpe_plan* p) {
    typedef typename sgpe::cx_of<T>::type C;
    std::vector<C> h((size_t)p->n1 * p->n2);
    for (long long k1 = 0; k1 < p->n1; k1++)
        for (long long q = 0; q < p->n2; q++) h[k1 * p->n2 + q] = unit_root<T>(k1 * q, (long long)p->n1 * p->n2);
    SGPE_CUDA(cudaMalloc(&p->tw4, sizeof(C) * h.size()));
    SGPE_CUDA(cudaMemcpy(p->tw4, h.data(), sizeof(C) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

template <typename T>
static int run_pack(sgpe_plan* p, const void* in, void* out, int A, int P, int Bw, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::PackArgs<T> a;
    a.in = static_cast<const C*>(in); a.out = static_cast<C*>(out); a.A = A; a.P = P; a.Bw = Bw;
    SGPE_LAUNCH((sgpe::slab_pack<T>), dim3(148 * 8), dim3(256), 0, st, a);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}
template <typename T>
static int run_unpack(sgpe_plan* p, const void* in, void* out, int P, int Bh, int Bw, cudaStream_t st) {
    typedef typename sgpe::cx_of<T>::type C;
    sgpe::UnpackArgs<T> a;
    a.in = static_cast<const C*>(in); a.out = static_cast<C*>(out); a.P = P; a.Bh = Bh; a.Bw = Bw;
    SGPE_LAUNCH((sgpe::slab_unpack_transpose<T>), dim3(148 * 8), dim3(256), 33 * 32 * sizeof(C), st, a);
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

// Energy of the step boundary whose un-normalised k-space state the junction pass just stored to `scratch` (and whose
// norm sums it left in `totals`): inverse transform on the side with ttools.norm folded into the last pass (device-side
// scale) and the per-component density maxima gathered by the same pass, then the stencil pass.  Three launches and
// no close / re-open of the junction, against six for sgpe_energy(NULL) between two sgpe_full_steps calls.
int energy_of_boundary(sgpe_plan* p, cudaStream_t st) {
    const double norm = p->dx * p->dy / (2.0 * M_PI);
    // (graph replay: node arguments are frozen, the slot comes from a device-side counter the stencil pass advances)
    int* const ectr = p->capturing ? p->eslot_ctr : nullptr;
    double* out = p->pend_energy + (ectr ? 0LL : 4LL * p->pend_eslot);
    // (the density maxima of the row pass are cleared by the column pass before it: no memset node)
    int rc = SGPE_BY_DTYPE(p, run_col, p, p->scratch, p->scratch, false, false, 0.0, false, 0.0, true, 0, 0, 1.0, nullptr,
                           0, -1, st, nullptr, p->maxdens);
    if (rc) return rc;
    // (option "energy_polar", default on: the last inverse pass stores (|psi|, arg psi) and the stencil pass is left
    // without square roots and arctangents; 0 = (re, im) through the tiled / streaming kernels as in round 1)
    const bool polar = p->energy_polar && p->energy_kernel == 0 && !p->generic;
    rc = SGPE_BY_DTYPE(p, run_row, p, p->scratch, p->scratch, true, false, 0.0, false, 0, 3,
                       1.0 / (norm * (double)p->nx * (double)p->ny), st, nullptr, 0.0, false, p->totals,
                       p->atom_num / p->dv_k, p->maxdens, polar);
    if (rc) return rc;
    rc = SGPE_BY_DTYPE(p, run_energy, p, p->scratch, p->track_unwrap, p->track_kl, out, st, p->pend_estride, true, polar, ectr);
    p->pend_eslot = -1;
    return rc;
}

}  // namespace

extern "C" {

const char* sgpe_last_error(void) { return g_err.c_str(); }

const char* sgpe_version(void) {
#ifdef SGPE_EMU
    return "sgpe 0.1 emu";
#else
    return "sgpe 0.1 sm_100a";
#endif
}

int sgpe_plan_create(sgpe_plan** out, int nx, int ny, int batch, int dtype, int device) {
    if (!out) return fail(SGPE_EINVAL, "null output pointer");
    *out = nullptr;
    const bool fused = sgpe::supported_length(nx) && sgpe::supported_length(ny);
    if (!fused && !(generic_length(nx) && generic_length(ny)))
        return fail(SGPE_EINVAL, "mesh sizes must be even, at most 4096 points per line, with prime factors 2, 3, 5, 7 "
                                 "(powers of two from 32 on run in the fused kernels; lines beyond 4096: sgpe_plan_create_lines)");
    if (batch < 1) return fail(SGPE_EINVAL, "batch must be >= 1");
    if (dtype != SGPE_C128 && dtype != SGPE_C64) return fail(SGPE_EINVAL, "dtype must be 0 (c128) or 1 (c64)");
    DeviceGuard guard(device);
    sgpe_plan* p = new sgpe_plan();
    p->nx = nx; p->ny = ny; p->batch = batch; p->dtype = dtype; p->device = device;
    p->generic = !fused;
    if (p->generic) { p->rad_x = generic_radices(nx); p->rad_y = generic_radices(ny); }
    p->csize = dtype == SGPE_C128 ? 16 : 8;
    p->plane = (long long)nx * ny;
    p->max_tiles = 2 * (nx > ny ? nx : ny);      // covers every column tile width and the line passes
    if (p->max_tiles < 1024) p->max_tiles = 1024;
    int rc = 0;
    do {
        if (cudaMalloc((void**)&p->partials, sizeof(double) * 2 * (size_t)p->max_tiles * batch) != cudaSuccess) { rc = SGPE_ENOMEM; break; }
        if (cudaMalloc((void**)&p->counter, sizeof(unsigned) * batch) != cudaSuccess) { rc = SGPE_ENOMEM; break; }
        if (cudaMalloc((void**)&p->totals, sizeof(double) * 4 * batch) != cudaSuccess) { rc = SGPE_ENOMEM; break; }
        if (cudaMalloc((void**)&p->totals_aux, sizeof(double) * 4 * batch) != cudaSuccess) { rc = SGPE_ENOMEM; break; }
        cudaMemset(p->counter, 0, sizeof(unsigned) * batch);
        cudaMemset(p->totals, 0, sizeof(double) * 4 * batch);
        cudaMemset(p->totals_aux, 0, sizeof(double) * 4 * batch);
        if (p->generic) break;                 // the generic transforms evaluate their twiddles in place
        rc = dtype == SGPE_C128 ? upload_twiddles<double>(&p->tw_x, nx) : upload_twiddles<float>(&p->tw_x, nx);
        if (rc) break;
        rc = dtype == SGPE_C128 ? upload_twiddles<double>(&p->tw_y, ny) : upload_twiddles<float>(&p->tw_y, ny);
    } while (0);
    if (rc) {
        if (rc == SGPE_ENOMEM) fail(rc, "device allocation failed");
        sgpe_plan_destroy(p);
        return rc;
    }
    *out = p;
    return 0;
}

int sgpe_plan_destroy(sgpe_plan* p) {
    if (!p) return 0;
    DeviceGuard guard(p->device);
    cudaFree(p->state); cudaFree(p->tw_x); cudaFree(p->tw_y); cudaFree(p->partials);
    cudaFree(p->counter); cudaFree(p->totals); cudaFree(p->totals_aux); cudaFree(p->pops_buf);
    cudaFree(p->scratch); cudaFree(p->maxdens); cudaFree(p->unwrap_inc); cudaFree(p->sm_slots); cudaFree(p->tw_mid); cudaFree(p->tw4);
    unwrap_release(p->unwrap);
    cudaFree(p->unwrap_phi);
    for (auto& t : p->tile_maps) delete t.map;
    cudaFree(p->slot_ctr);
    cudaFree(p->eslot_ctr);
#ifndef SGPE_EMU
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    if (p->graph_exec_e) cudaGraphExecDestroy(p->graph_exec_e);
    if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
#endif
    for (auto& t : p->kin_tab) { cudaFree(t.x); cudaFree(t.y); }
    for (auto& t : p->pot_tab) { cudaFree(t.x); cudaFree(t.y); }
    delete p;
    return 0;
}

int sgpe_set_grid(sgpe_plan* p, double dx, double dy, double dv_r, double dv_k, double atom_num) {
    if (p) p->epoch++;
    if (!p) return fail(SGPE_EINVAL, "null plan");
    if (!(dx > 0 && dy > 0 && dv_r > 0 && dv_k > 0 && atom_num > 0)) return fail(SGPE_EINVAL, "grid values must be positive");
    p->dx = dx; p->dy = dy; p->dv_r = dv_r; p->dv_k = dv_k; p->atom_num = atom_num; p->grid_set = true;
    return 0;
}

int sgpe_set_interactions(sgpe_plan* p, double g_uu, double g_dd, double g_ud) {
    if (p) p->epoch++;
    if (!p) return fail(SGPE_EINVAL, "null plan");
    p->g_uu = g_uu; p->g_dd = g_dd; p->g_ud = g_ud; p->g_set = true;
    return 0;
}

int sgpe_set_kinetic(sgpe_plan* p, const double* kin0, const double* kin1, int64_t bs) {
    if (p) p->epoch++;
    if (!p || !kin0 || !kin1) return fail(SGPE_EINVAL, "null argument");
    p->kin0 = kin0; p->kin1 = kin1; p->kin_bs = bs; p->kin_set = true; p->kin_mode = 0;
    return 0;
}

int sgpe_set_potential(sgpe_plan* p, const double* pot0, const double* pot1, int64_t bs) {
    if (p) p->epoch++;
    if (!p || !pot0 || !pot1) return fail(SGPE_EINVAL, "null argument");
    p->pot0 = pot0; p->pot1 = pot1; p->pot_bs = bs; p->pot_set = true; p->pot_mode = 0;
    return 0;
}

int sgpe_set_kinetic_separable(sgpe_plan* p, const double* kin_x, const double* kin_y, int64_t xbs, int64_t ybs) {
    if (p) p->epoch++;
    if (!p || !kin_x || !kin_y) return fail(SGPE_EINVAL, "null argument");
    p->kin_x = kin_x; p->kin_y = kin_y; p->kin_xbs = xbs; p->kin_ybs = ybs; p->kin_mode = 1; p->kin_set = true;
    invalidate_tables(p->kin_tab, 6);
    return 0;
}

int sgpe_set_potential_separable(sgpe_plan* p, const double* pot_x, const double* pot_y, int64_t xbs, int64_t ybs) {
    if (p) p->epoch++;
    if (!p || !pot_x || !pot_y) return fail(SGPE_EINVAL, "null argument");
    p->pot_x = pot_x; p->pot_y = pot_y; p->pot_xbs = xbs; p->pot_ybs = ybs; p->pot_mode = 1; p->pot_set = true;
    invalidate_tables(p->pot_tab, 4);
    return 0;
}

int sgpe_set_coupling(sgpe_plan* p, int mode, const double* coupling, int64_t bs, const double* omega,
                      const void* eiphi) {
    if (p) p->epoch++;
    if (!p) return fail(SGPE_EINVAL, "null plan");
    if (mode == SGPE_COUPLING_UNIFORM && !omega) return fail(SGPE_EINVAL, "uniform coupling needs omega_dev[batch]");
    if (mode == SGPE_COUPLING_DENSE && !coupling) return fail(SGPE_EINVAL, "dense coupling needs coupling_dev");
    if (mode < 0 || mode > 2) return fail(SGPE_EINVAL, "bad coupling mode");
    p->cpl_mode = mode; p->cpl = coupling; p->cpl_bs = bs; p->omega = omega; p->eiphi = eiphi;
    return 0;
}

int sgpe_set_energy_coupling(sgpe_plan* p, int mode, const double* coupling, int64_t bs, const double* omega) {
    if (p) p->epoch++;
    if (!p) return fail(SGPE_EINVAL, "null plan");
    if (mode < -1 || mode > 2) return fail(SGPE_EINVAL, "bad coupling mode");
    if (mode == SGPE_COUPLING_UNIFORM && !omega) return fail(SGPE_EINVAL, "uniform coupling needs omega_dev[batch]");
    if (mode == SGPE_COUPLING_DENSE && !coupling) return fail(SGPE_EINVAL, "dense coupling needs coupling_dev");
    p->ecpl_mode = mode; p->ecpl = coupling; p->ecpl_bs = bs; p->eomega = omega;
    return 0;
}

int sgpe_set_option(sgpe_plan* p, const char* name, int value) {
    if (p) p->epoch++;
    if (!p || !name) return fail(SGPE_EINVAL, "null argument");
    if (std::strcmp(name, "col_tile") == 0) {
        if (value != 0 && value != 2 && value != 3 && value != 8)
            return fail(SGPE_EINVAL, "col_tile: 0 (default), 2 (half width), 3 (two barrier groups) or 8 (radix-8 threads)");
        p->col_wsel = value;
        return 0;
    }
    if (std::strcmp(name, "stagger_ns") == 0) { p->stagger_ns = value; return 0; }
    if (std::strcmp(name, "timeline_kind") == 0) { p->dbg_kind = value ? 1 : 0; return 0; }
    if (std::strcmp(name, "energy_kernel") == 0) { p->energy_kernel = value ? 1 : 0; return 0; }
    if (std::strcmp(name, "energy_polar") == 0) {     // 0: (re, im); 1: polar + warp-per-band stencils; 2: polar + streaming kernel
        if (value < 0 || value > 2) return fail(SGPE_EINVAL, "energy_polar: 0, 1 or 2");
        p->energy_polar = value;
        return 0;
    }
    if (std::strcmp(name, "graph") == 0) {
        if (value < -1 || value > 1) return fail(SGPE_EINVAL, "graph: -1 (default: on), 0 (off) or 1 (on)");
        p->use_graph = value;
        return 0;
    }
    if (std::strcmp(name, "col_kernel") == 0) {
        if (value < 0 || value > 7)
            return fail(SGPE_EINVAL, "col_kernel: 0 (default), 1 (one tile per CTA), 2 (persistent, TMA-staged, split inverse "
                                     "exchange), 3 (persistent, TMA-staged behind the inverse transform), 4 (persistent, "
                                     "two barrier groups per CTA) 5 (as 2 with the twiddle tables in shared memory), 6 (as 2 with half-width tiles, two CTAs per SM) or 7 (as 2 "
                                     "with the k factors of the launch in shared memory, imaginary time)");
        p->col_kernel = value;
        return 0;
    }
    if (std::strcmp(name, "prefetch") == 0) { p->prefetch = value ? 1 : 0; return 0; }
    if (std::strcmp(name, "unwrap_merge") == 0) {
        if (value != 0 && value != 1) return fail(SGPE_EINVAL, "unwrap_merge: 0 (spanning tree on the device) or 1 (all on the host)");
        p->unwrap_merge = value;
        return 0;
    }
    if (std::strcmp(name, "unwrap_tail") == 0) {
        if (value < 0) return fail(SGPE_EINVAL, "unwrap_tail: a non-negative number of edges");
        p->unwrap_tail = value;
        return 0;
    }
    if (std::strcmp(name, "unwrap_anchor") == 0) {
        if (value < -1 || value > 1) return fail(SGPE_EINVAL, "unwrap_anchor: -1 (by plane count), 0 (host pass) or 1 (device bisection)");
        p->unwrap_anchor = value;
        return 0;
    }
    if (std::strcmp(name, "unwrap_sort") == 0) {
        if (value != 0 && value != 1) return fail(SGPE_EINVAL, "unwrap_sort: 0 (device radix sort) or 1 (host sort)");
        p->unwrap_sort = value;
        return 0;
    }
    if (std::strcmp(name, "row_mode") == 0) {
        if (value != 0 && value != 1) return fail(SGPE_EINVAL, "row_mode: 0 (paired) or 1 (split)");
        p->row_mode = value;
        return 0;
    }
    return fail(SGPE_EINVAL, std::string("unknown option ") + name);
}

int sgpe_set_time(sgpe_plan* p, int time_mode, double dt) {
    if (p) p->epoch++;
    if (!p) return fail(SGPE_EINVAL, "null plan");
    if (time_mode != SGPE_TIME_REAL && time_mode != SGPE_TIME_IMAG) return fail(SGPE_EINVAL, "bad time mode");
    if (p->phase == sgpe_plan::MID && time_mode != p->tm)
        return fail(SGPE_ESTATE, "store the state before switching between real and imaginary time");
    p->tm = time_mode; p->dt = dt;
    p->dt_out = dt * kGamma;                 // tensor_propagator.py:102
    p->dt_in = dt * (1.0 - 2.0 * kGamma);    // tensor_propagator.py:103
    p->time_set = true;
    return 0;
}

int sgpe_substeps(const sgpe_plan* p, double* dt_out, double* dt_in) {
    if (!p || !p->time_set) return fail(SGPE_ESTATE, "time not set");
    if (dt_out) *dt_out = p->dt_out;
    if (dt_in) *dt_in = p->dt_in;
    return 0;
}

int sgpe_load_psik(sgpe_plan* p, const void* psik, sgpe_stream st) {
    if (!p || !psik) return fail(SGPE_EINVAL, "null argument");
    DeviceGuard guard(p->device);
    if (int rc = ensure_state(p)) return rc;
    SGPE_CUDA(cudaMemcpyAsync(p->state, psik, (size_t)p->batch * 2 * p->plane * p->csize, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)st));
    p->phase = sgpe_plan::KSPACE; p->scale_pending = false; p->pend_slot = -1; p->pend_eslot = -1;
    return 0;
}

int sgpe_store_psik(sgpe_plan* p, void* psik, sgpe_stream st) {
    if (!p || !psik) return fail(SGPE_EINVAL, "null argument");
    if (p->phase == sgpe_plan::EMPTY) return fail(SGPE_ESTATE, "no state loaded");
    DeviceGuard guard(p->device);
    int rc = close_junction(p, (cudaStream_t)st);
    if (rc) return rc;
    if (p->scale_pending) {
        rc = SGPE_BY_DTYPE(p, run_scale, p, p->state, psik, p->totals, p->atom_num / p->dv_k, (cudaStream_t)st);
        if (rc) return rc;
        if (psik == p->state) p->scale_pending = false;
    } else if (psik != p->state) {
        SGPE_CUDA(cudaMemcpyAsync(psik, p->state, (size_t)p->batch * 2 * p->plane * p->csize,
                                  cudaMemcpyDeviceToDevice, (cudaStream_t)st));
    }
    return 0;
}

int sgpe_single_step(sgpe_plan* p, double dt_sub, sgpe_stream st) {
    int rc = ready_to_step(p);
    if (rc) return rc;
    DeviceGuard guard(p->device);
    return single_step_impl(p, dt_sub, (cudaStream_t)st);
}

#ifndef SGPE_EMU
// One steady-state full step (MID -> MID, the junction that opens it records the populations of the step before)
// captured into an executable graph.  Called with every factor table the step needs already cached.
static int capture_step_graph(sgpe_plan* p, double* pops, int64_t pops_stride, double* energy = nullptr, int64_t energy_stride = 0) {
    // (captured on a stream of the plan's own: the caller's may be the legacy default stream, which cannot capture;
    // the executable graph is then launched on the caller's stream)
    if (!p->cap_stream) SGPE_CUDA(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
    cudaStream_t s = p->cap_stream;
    cudaGraphExec_t& exec = energy ? p->graph_exec_e : p->graph_exec;
    if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
    if (!p->slot_ctr && cudaMalloc((void**)&p->slot_ctr, sizeof(int) * p->batch) != cudaSuccess)
        return fail(SGPE_ENOMEM, "device allocation failed");
    if (!p->eslot_ctr && cudaMalloc((void**)&p->eslot_ctr, sizeof(int) * p->batch) != cudaSuccess)
        return fail(SGPE_ENOMEM, "device allocation failed");
    cudaGraph_t graph = nullptr;
    SGPE_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    p->capturing = true;
    p->pend_pops = pops; p->pend_stride = pops_stride; p->pend_slot = pops ? 0 : -1;     // slot value unused: counter
    p->pend_energy = energy; p->pend_estride = energy_stride; p->pend_eslot = energy ? 0 : -1;
    int rc = single_step_impl(p, p->dt_out, s);
    if (!rc) rc = single_step_impl(p, p->dt_in, s);
    if (!rc) rc = single_step_impl(p, p->dt_out, s);
    p->capturing = false;
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(SGPE_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { exec = nullptr; return fail(SGPE_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    return 0;
}
#endif

// should sgpe_full_steps replay a captured graph?  Default: yes from four steps on (launch-bound small meshes gain 20 %,
// 2048^2 still 1.6 %: 1596 -> 1622 steps/s); long-line plans, profiling runs and streams that are already being captured
// never do.
static bool graph_wanted(const sgpe_plan* p, int n, cudaStream_t s) {
#ifdef SGPE_EMU
    (void)p; (void)n; (void)s;
    return false;
#else
    if (p->use_graph == 0 || n < 4 || p->prof_on || p->n1 != 1 || p->dbg || p->dbg_col) return false;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return false; }
    return true;
#endif
}

int sgpe_full_steps(sgpe_plan* p, int n, double* pops, int64_t pops_stride, int pops_first, sgpe_stream st) {
    int rc = ready_to_step(p);
    if (rc) return rc;
    if (n < 0) return fail(SGPE_EINVAL, "negative step count");
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)st;
#ifndef SGPE_EMU
    if (graph_wanted(p, n, s)) {
        // steps 0 and 1 launch normally (they open the junction and leave every factor table of the steady state in
        // the cache), steps 2 .. n-1 replay the captured step, the last junction is closed normally
        for (int i = 0; i < 2; i++) {
            if ((rc = single_step_impl(p, p->dt_out, s))) return rc;
            if ((rc = single_step_impl(p, p->dt_in, s))) return rc;
            if ((rc = single_step_impl(p, p->dt_out, s))) return rc;
            p->pend_pops = pops; p->pend_stride = pops_stride; p->pend_slot = pops ? pops_first + i : -1;
        }
        sgpe_plan::GraphKey key;
        key.epoch = p->epoch; key.pops = pops; key.stride = pops_stride; key.tm = p->tm; key.dt = p->dt;
        if (!p->graph_exec || !(p->graph_key == key)) {
            // (the capture itself executes nothing: the state machine is restored afterwards)
            const sgpe_plan::Phase phase = p->phase; const double pend = p->pending_dt; const bool sc = p->scale_pending;
            const uint64_t l0 = p->launches;
            rc = capture_step_graph(p, pops, pops_stride);
            p->phase = phase; p->pending_dt = pend; p->scale_pending = sc; p->launches = l0;
            if (rc) return rc;
            p->graph_key = key;
        }
        if (pops) {
            sgpe::fill_value<int><<<(p->batch + 127) / 128, 128, 0, s>>>(p->slot_ctr, pops_first + 1, p->batch);
            p->launches++;
        }
        for (int i = 2; i < n; i++) SGPE_CUDA(cudaGraphLaunch(p->graph_exec, s));
        p->launches += 6ull * (uint64_t)(n - 2);
        p->phase = sgpe_plan::MID; p->pending_dt = p->dt_out; p->scale_pending = false;
        p->pend_pops = pops; p->pend_stride = pops_stride; p->pend_slot = pops ? pops_first + n - 1 : -1;
        p->pend_eslot = -1;
        return close_junction(p, s);
    }
#endif
    for (int i = 0; i < n; i++) {                       // tensor_propagator.py:220-222
        if ((rc = single_step_impl(p, p->dt_out, s))) return rc;
        if ((rc = single_step_impl(p, p->dt_in, s))) return rc;
        if ((rc = single_step_impl(p, p->dt_out, s))) return rc;
        p->pend_pops = pops; p->pend_stride = pops_stride; p->pend_slot = pops ? pops_first + i : -1;
    }
    // close the last junction so that the populations of the final step are recorded before returning
    if (n > 0 && (rc = close_junction(p, s))) return rc;
    return 0;
}

static int energy_of_real_space(sgpe_plan* p, const void* psi, int unwrap_mode, double kl_term, double* out, sgpe_stream st,
                                long long out_bstride);

int sgpe_full_steps_energy(sgpe_plan* p, int n, double* pops, int64_t pops_stride, int pops_first, double* energy,
                           int64_t energy_stride, int energy_first, int unwrap_mode, double kl_term, sgpe_stream st) {
    int rc = ready_to_step(p);
    if (rc) return rc;
    if (n < 0) return fail(SGPE_EINVAL, "negative step count");
    if (!energy) return fail(SGPE_EINVAL, "null energy buffer (use sgpe_full_steps)");
    if (unwrap_mode < 0 || unwrap_mode > 2) return fail(SGPE_EINVAL, "unwrap_mode must be 0, 1 or 2");
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)st;
    const size_t bytes = (size_t)p->batch * 2 * p->plane * p->csize;
    if (!p->scratch && cudaMalloc(&p->scratch, bytes) != cudaSuccess) return fail(SGPE_ENOMEM, "scratch allocation failed");
    if (!p->maxdens && cudaMalloc((void**)&p->maxdens, sizeof(double) * 2 * p->batch) != cudaSuccess)
        return fail(SGPE_ENOMEM, "scratch allocation failed");
    if (unwrap_mode == 2) {
        // the reference's own definition (phase unwrapped by reliability-sorted region merging) after every step: the
        // stand-alone evaluation of the closed step, steered from the host (synchronises the stream once per step)
        for (int i = 0; i < n; i++) {
            if ((rc = sgpe_full_steps(p, 1, pops, pops_stride, pops_first + i, st))) return rc;
            if ((rc = sgpe_store_psik(p, p->scratch, st))) return rc;
            if ((rc = sgpe_fft2d(p, p->scratch, p->scratch, 1, st))) return rc;
            if ((rc = energy_of_real_space(p, p->scratch, 2, kl_term, energy + 4LL * (energy_first + i), st, energy_stride))) return rc;
        }
        return 0;
    }
    if (p->generic) {
        // generic meshes: the energy of every step through the stand-alone evaluation (junction closed each step)
        for (int i = 0; i < n; i++) {
            if ((rc = sgpe_full_steps(p, 1, pops, pops_stride, pops_first + i, st))) return rc;
            if ((rc = sgpe_store_psik(p, p->scratch, st))) return rc;
            if ((rc = sgpe_fft2d(p, p->scratch, p->scratch, 1, st))) return rc;
            if ((rc = SGPE_BY_DTYPE(p, run_energy, p, p->scratch, unwrap_mode, kl_term, energy + 4LL * (energy_first + i), s,
                                    energy_stride)))
                return rc;
        }
        return 0;
    }
    p->track_unwrap = unwrap_mode; p->track_kl = kl_term;
#ifndef SGPE_EMU
    if (graph_wanted(p, n, s) && p->energy_kernel == 0) {
        // as in sgpe_full_steps: two steps launched normally, the steady-state step (junction with the side store, the
        // three passes of the energy chain, five more passes) replayed n - 2 times, the last junction closed normally
        for (int i = 0; i < 2; i++) {
            if ((rc = single_step_impl(p, p->dt_out, s))) return rc;
            if ((rc = single_step_impl(p, p->dt_in, s))) return rc;
            if ((rc = single_step_impl(p, p->dt_out, s))) return rc;
            p->pend_pops = pops; p->pend_stride = pops_stride; p->pend_slot = pops ? pops_first + i : -1;
            p->pend_energy = energy; p->pend_estride = energy_stride; p->pend_eslot = energy_first + i;
        }
        sgpe_plan::GraphKey key;
        key.epoch = p->epoch; key.pops = pops; key.stride = pops_stride; key.tm = p->tm; key.dt = p->dt;
        key.energy = energy; key.estride = energy_stride; key.unwrap = unwrap_mode; key.kl = kl_term;
        if (!p->graph_exec_e || !(p->graph_key_e == key)) {
            const sgpe_plan::Phase phase = p->phase; const double pend = p->pending_dt; const bool sc = p->scale_pending;
            const uint64_t l0 = p->launches;
            rc = capture_step_graph(p, pops, pops_stride, energy, energy_stride);
            p->phase = phase; p->pending_dt = pend; p->scale_pending = sc; p->launches = l0;
            if (rc) return rc;
            p->graph_key_e = key;
        }
        if (pops) sgpe::fill_value<int><<<(p->batch + 127) / 128, 128, 0, s>>>(p->slot_ctr, pops_first + 1, p->batch);
        sgpe::fill_value<int><<<(p->batch + 127) / 128, 128, 0, s>>>(p->eslot_ctr, energy_first + 1, p->batch);
        p->launches += pops ? 2 : 1;
        for (int i = 2; i < n; i++) SGPE_CUDA(cudaGraphLaunch(p->graph_exec_e, s));
        p->launches += 9ull * (uint64_t)(n - 2);
        p->phase = sgpe_plan::MID; p->pending_dt = p->dt_out; p->scale_pending = false;
        p->pend_pops = pops; p->pend_stride = pops_stride; p->pend_slot = pops ? pops_first + n - 1 : -1;
        p->pend_energy = energy; p->pend_estride = energy_stride; p->pend_eslot = energy_first + n - 1;
        return close_junction(p, s);
    }
#endif
    for (int i = 0; i < n; i++) {
        if ((rc = single_step_impl(p, p->dt_out, s))) return rc;
        if ((rc = single_step_impl(p, p->dt_in, s))) return rc;
        if ((rc = single_step_impl(p, p->dt_out, s))) return rc;
        p->pend_pops = pops; p->pend_stride = pops_stride; p->pend_slot = pops ? pops_first + i : -1;
        p->pend_energy = energy; p->pend_estride = energy_stride; p->pend_eslot = energy_first + i;
    }
    if (n > 0 && (rc = close_junction(p, s))) return rc;
    return 0;
}

int sgpe_fft2d(sgpe_plan* p, const void* in, void* out, int inverse, sgpe_stream st) {
    if (!p || !in || !out) return fail(SGPE_EINVAL, "null argument");
    if (!p->grid_set) return fail(SGPE_ESTATE, "grid not set");
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)st;
    const double norm = p->dx * p->dy / (2.0 * M_PI);     // tensor_tools.py:218, 248
    int rc;
    if (!inverse) {
        rc = SGPE_BY_DTYPE(p, run_row, p, in, out, false, false, 0.0, true, 3, 0, norm, s);
        if (rc) return rc;
        return SGPE_BY_DTYPE(p, run_col, p, out, out, true, false, 0.0, false, 0.0, false, 0, 0, 1.0, nullptr, 0, -1, s);
    }
    rc = SGPE_BY_DTYPE(p, run_col, p, in, out, false, false, 0.0, false, 0.0, true, 0, 0, 1.0, nullptr, 0, -1, s);
    if (rc) return rc;
    return SGPE_BY_DTYPE(p, run_row, p, out, out, true, false, 0.0, false, 0, 3,
                         1.0 / (norm * (double)p->nx * (double)p->ny), s);
}

int sgpe_fft1d(sgpe_plan* p, const void* in, void* out, int axis, int inverse, sgpe_stream st) {
    if (!p || !in || !out) return fail(SGPE_EINVAL, "null argument");
    if (!p->grid_set) return fail(SGPE_ESTATE, "grid not set");
    if (axis != 0 && axis != 1) return fail(SGPE_EINVAL, "axis must be 0 (x) or 1 (y)");
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)st;
    const double norm = (axis == 0 ? p->dx : p->dy) / std::sqrt(2.0 * M_PI);   // tensor_tools.py:151, 188
    const int n = axis == 0 ? p->nx : p->ny;
    if (axis == 0) {
        if (!inverse) return SGPE_BY_DTYPE(p, run_row, p, in, out, false, false, 0.0, true, 1, 0, norm, s);
        return SGPE_BY_DTYPE(p, run_row, p, in, out, true, false, 0.0, false, 0, 1, 1.0 / (norm * n), s);
    }
    if (!inverse)
        return SGPE_BY_DTYPE(p, run_col, p, in, out, true, false, 0.0, false, 0.0, false, 1, 0, norm, nullptr, 0, -1, s);
    return SGPE_BY_DTYPE(p, run_col, p, in, out, false, false, 0.0, false, 0.0, true, 0, 1, 1.0 / (norm * n), nullptr,
                         0, -1, s);
}

int sgpe_sumsq(sgpe_plan* p, const void* in, double* out, sgpe_stream st) {
    if (!p || !in || !out) return fail(SGPE_EINVAL, "null argument");
    DeviceGuard guard(p->device);
    return SGPE_BY_DTYPE(p, run_sumsq, p, in, out, (cudaStream_t)st);
}

int sgpe_normalise(sgpe_plan* p, const void* in, void* out, double vol, sgpe_stream st) {
    if (!p || !in || !out) return fail(SGPE_EINVAL, "null argument");
    if (!p->grid_set) return fail(SGPE_ESTATE, "grid not set (atom number)");
    DeviceGuard guard(p->device);
    int rc = SGPE_BY_DTYPE(p, run_sumsq, p, in, nullptr, (cudaStream_t)st);
    if (rc) return rc;
    return SGPE_BY_DTYPE(p, run_scale, p, in, out, p->totals_aux, p->atom_num / vol, (cudaStream_t)st);
}

// the energy functional on a real-space state (the part of eng_expect after its ifft_2d, tensor_propagator.py:301-324)
static int energy_of_real_space(sgpe_plan* p, const void* psi, int unwrap_mode, double kl_term, double* out, sgpe_stream st,
                                long long out_bstride) {
    if (!p->maxdens && cudaMalloc((void**)&p->maxdens, sizeof(double) * 2 * p->batch) != cudaSuccess)
        return fail(SGPE_ENOMEM, "scratch allocation failed");
    if (unwrap_mode == 2) {
        // wrapped phase of the real-space state -> integer field of 2 pi multiples (synchronises the stream)
        const long long total = (long long)p->batch * 2 * p->plane;
        if (!p->unwrap_inc && cudaMalloc((void**)&p->unwrap_inc, sizeof(int) * total) != cudaSuccess)
            return fail(SGPE_ENOMEM, "unwrap allocation failed");
        if (!p->unwrap_phi && cudaMalloc((void**)&p->unwrap_phi, sizeof(double) * total) != cudaSuccess)
            return fail(SGPE_ENOMEM, "unwrap allocation failed");
        int rc = SGPE_BY_DTYPE(p, run_unwrap_angles, p, psi, total, p->unwrap_phi, (cudaStream_t)st);
        if (!rc) rc = unwrap_increments(p, p->unwrap_phi, 2 * p->batch, p->unwrap_inc, (cudaStream_t)st);
        if (rc) return rc;
    }
    return SGPE_BY_DTYPE(p, run_energy, p, psi, unwrap_mode, kl_term, out, (cudaStream_t)st, out_bstride);
}

int sgpe_energy(sgpe_plan* p, const void* psik, int unwrap_mode, double kl_term, double* out, sgpe_stream st) {
    if (!p || !out) return fail(SGPE_EINVAL, "null argument");
    if (!(p->grid_set && p->g_set && p->pot_set)) return fail(SGPE_ESTATE, "set grid, interactions and potential first");
    if (unwrap_mode < 0 || unwrap_mode > 2) return fail(SGPE_EINVAL, "unwrap_mode must be 0, 1 or 2");
    if (p->n1 != 1 || p->line_plan) return fail(SGPE_EINVAL, "line plans have no k-space state: use sgpe_energy_real_space");
    DeviceGuard guard(p->device);
    const size_t bytes = (size_t)p->batch * 2 * p->plane * p->csize;
    if (!p->scratch && cudaMalloc(&p->scratch, bytes) != cudaSuccess) return fail(SGPE_ENOMEM, "scratch allocation failed");
    int rc;
    if (psik == nullptr) {
        if (p->phase == sgpe_plan::EMPTY) return fail(SGPE_ESTATE, "no state loaded");
        if ((rc = sgpe_store_psik(p, p->scratch, st))) return rc;
        psik = p->scratch;
    }
    if ((rc = sgpe_fft2d(p, psik, p->scratch, 1, st))) return rc;
    return energy_of_real_space(p, p->scratch, unwrap_mode, kl_term, out, st, 4);
}

int sgpe_kinetic_spectral(sgpe_plan* p, const void* psik, double* out, sgpe_stream st) {
    if (!p || !out) return fail(SGPE_EINVAL, "null argument");
    if (!(p->grid_set && p->kin_set)) return fail(SGPE_ESTATE, "set grid and kinetic operator first");
    if (p->n1 != 1 || p->line_plan) return fail(SGPE_EINVAL, "not available on line plans");
    DeviceGuard guard(p->device);
    if (psik == nullptr) {
        if (p->phase == sgpe_plan::EMPTY) return fail(SGPE_ESTATE, "no state loaded");
        const size_t bytes = (size_t)p->batch * 2 * p->plane * p->csize;
        if (!p->scratch && cudaMalloc(&p->scratch, bytes) != cudaSuccess) return fail(SGPE_ENOMEM, "scratch allocation failed");
        int rc = sgpe_store_psik(p, p->scratch, st);
        if (rc) return rc;
        psik = p->scratch;
    }
    return SGPE_BY_DTYPE(p, run_kinetic, p, psik, out, (cudaStream_t)st);
}

int sgpe_gradient(sgpe_plan* p, const void* f, int is_complex, double h0, double h1, void* g0, void* g1, sgpe_stream st) {
    if (!p || !f || !g0 || !g1) return fail(SGPE_EINVAL, "null argument");
    if (p->nx < 2 || p->ny < 2) return fail(SGPE_EINVAL, "np.gradient needs at least two points per axis");
    if (!(h0 != 0.0) || !(h1 != 0.0)) return fail(SGPE_EINVAL, "zero spacing");
    DeviceGuard guard(p->device);
    const int nc = is_complex ? 2 : 1;
    const long long n = (long long)p->plane * nc;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (p->dtype == SGPE_C128) {
        SGPE_LAUNCH((sgpe::gradient_pass<double>), dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)st,
                    static_cast<const double*>(f), p->ny, p->nx, nc, 1.0 / h0, 1.0 / h1, static_cast<double*>(g0), static_cast<double*>(g1));
    } else {
        SGPE_LAUNCH((sgpe::gradient_pass<float>), dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)st,
                    static_cast<const float*>(f), p->ny, p->nx, nc, 1.0 / h0, 1.0 / h1, static_cast<float*>(g0), static_cast<float*>(g1));
    }
    p->launches++;
    SGPE_CUDA(cudaGetLastError());
    return 0;
}

int sgpe_energy_real_space(sgpe_plan* p, const void* psi, int unwrap_mode, double kl_term, double* out, sgpe_stream st) {
    if (!p || !psi || !out) return fail(SGPE_EINVAL, "null argument");
    if (!(p->grid_set && p->g_set && p->pot_set)) return fail(SGPE_ESTATE, "set grid, interactions and potential first");
    if (unwrap_mode < 0 || unwrap_mode > 2) return fail(SGPE_EINVAL, "unwrap_mode must be 0, 1 or 2");
    DeviceGuard guard(p->device);
    return energy_of_real_space(p, psi, unwrap_mode, kl_term, out, st, 4);
}

int sgpe_unwrap_phase(sgpe_plan* p, const void* in, int kind, int nplanes, int mask, double* out, sgpe_stream st_) {
    if (!p || !in || !out) return fail(SGPE_EINVAL, "null argument");
    if (kind != 0 && kind != 1) return fail(SGPE_EINVAL, "kind must be 0 (complex field) or 1 (wrapped angles)");
    if (nplanes < 1) return fail(SGPE_EINVAL, "nplanes must be >= 1");
    if (mask && kind != 0) return fail(SGPE_EINVAL, "the density mask needs the complex field (kind 0)");
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)st_;
    const long long total = (long long)nplanes * p->plane;
    struct Scratch { double* phi = nullptr; int* inc = nullptr; unsigned long long* maxbits = nullptr;
                     ~Scratch() { cudaFree(phi); cudaFree(inc); cudaFree(maxbits); } } w;
    if (cudaMalloc((void**)&w.inc, sizeof(int) * total) != cudaSuccess ||
        cudaMalloc((void**)&w.maxbits, sizeof(unsigned long long) * nplanes) != cudaSuccess ||
        (kind == 0 && cudaMalloc((void**)&w.phi, sizeof(double) * total) != cudaSuccess))
        return fail(SGPE_ENOMEM, "unwrap allocation failed");
    int rc;
    const double* phi = static_cast<const double*>(in);
    if (kind == 0) {
        if ((rc = SGPE_BY_DTYPE(p, run_unwrap_angles, p, in, total, w.phi, st))) return rc;
        phi = w.phi;
    }
    if ((rc = unwrap_increments(p, phi, nplanes, w.inc, st))) return rc;
    if ((rc = SGPE_BY_DTYPE(p, run_unwrap_apply, p, in, phi, w.inc, nplanes, mask != 0, w.maxbits, out, st))) return rc;
    SGPE_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---- slab (distributed) mode building blocks: the caller (spinor_gpe_b200/slab.py) owns the buffers and the
// collectives (torch.distributed all_to_all / all_reduce over NCCL); these are the local passes.
static int scatter_ready(const sgpe_plan* p, int scatter, int mode) {
    if (!scatter) return 0;
    if (p->peers.n < 1) return fail(SGPE_ESTATE, "scatter store requested before sgpe_slab_set_peers");
    if (p->peers.mode != mode) return fail(SGPE_EINVAL, "this pass cannot store through the plan's exchange map");
    return 0;
}

int sgpe_pass_rows(sgpe_plan* p, void* buf, double dt_sub, const double* totals_dev, double global_points,
                   int scatter, sgpe_stream st) {
    if (!p || !buf || !totals_dev) return fail(SGPE_EINVAL, "null argument");
    if (!(p->grid_set && p->g_set && p->pot_set && p->time_set)) return fail(SGPE_ESTATE, "plan not configured");
    int rc = scatter_ready(p, scatter, 1);
    if (rc) return rc;
    DeviceGuard guard(p->device);
    return SGPE_BY_DTYPE(p, run_row, p, buf, buf, true, true, dt_sub, true, 0, 0, 1.0, (cudaStream_t)st, totals_dev,
                         global_points, scatter != 0);
}

int sgpe_slab_set_peers(sgpe_plan* p, void* const* peer_bufs, int nranks, int mode, int seg, int drow, int64_t dplane,
                        int base) {
    if (p) p->epoch++;
    if (!p || !peer_bufs) return fail(SGPE_EINVAL, "null argument");
    if (nranks < 1 || nranks > SGPE_MAX_PEERS) return fail(SGPE_EINVAL, "1..16 ranks");
    if (mode != 1 && mode != 2) return fail(SGPE_EINVAL, "mode must be 1 (split along x) or 2 (split along y)");
    if (seg < 1 || drow < 1 || dplane < 1 || base < 0) return fail(SGPE_EINVAL, "bad exchange geometry");
    for (int i = 0; i < nranks; i++) {
        if (!peer_bufs[i]) return fail(SGPE_EINVAL, "null peer buffer");
        p->peers.ptr[i] = peer_bufs[i];
    }
    p->peers.n = nranks; p->peers.mode = mode; p->peers.seg = seg; p->peers.drow = drow; p->peers.dplane = dplane;
    p->peers.base = base;
    return 0;
}

int sgpe_slab_window(sgpe_plan* p, int first, int count, int chunk, int max_ctas) {
    if (!p) return fail(SGPE_EINVAL, "null plan");
    if (count == 0) { p->win = sgpe_plan::Window(); return 0; }
    if (first < 0 || count < 0 || first + count > p->ny) return fail(SGPE_EINVAL, "window outside the slab");
    if (chunk < 0 || chunk >= SGPE_MAX_CHUNKS) return fail(SGPE_EINVAL, "reduction slot out of range");
    if ((long long)(chunk + 1) * count > p->ny) return fail(SGPE_EINVAL, "reduction slots assume equal windows: (chunk + 1) * count <= lines");
    if (max_ctas < 0) return fail(SGPE_EINVAL, "negative grid cap");
    p->win.first = first; p->win.count = count; p->win.chunk = chunk; p->win.max_ctas = max_ctas;
    return 0;
}

int sgpe_pass_kcols(sgpe_plan* p, void* buf, int do_fwd, int has_a, double tau_a, int has_b, double tau_b, int do_inv,
                    double* sums_dev, int scatter, sgpe_stream st) {
    if (!p || !buf) return fail(SGPE_EINVAL, "null argument");
    if ((has_a || has_b) && (!p->kin_set || !p->time_set || !sums_dev)) return fail(SGPE_ESTATE, "kinetic operator / time / sums missing");
    if (p->batch != 1) return fail(SGPE_EINVAL, "line passes are for batch == 1 plans");
    int rc = scatter_ready(p, scatter, 2);
    if (rc) return rc;
    DeviceGuard guard(p->device);
    return SGPE_BY_DTYPE(p, run_kcols, p, buf, do_fwd != 0, has_a != 0, tau_a, has_b != 0, tau_b, do_inv != 0,
                         sums_dev, scatter != 0, (cudaStream_t)st);
}

// ---- exchange buffers visible to the other ranks of the node (one process per GPU): cudaMalloc + CUDA IPC.
int sgpe_ipc_alloc(int device, uint64_t bytes, void** dev_ptr, unsigned char handle[SGPE_IPC_HANDLE_BYTES]) {
    if (!dev_ptr || !handle || bytes == 0) return fail(SGPE_EINVAL, "bad argument");
    *dev_ptr = nullptr;
    memset(handle, 0, SGPE_IPC_HANDLE_BYTES);
#ifndef SGPE_EMU
    static_assert(sizeof(cudaIpcMemHandle_t) <= SGPE_IPC_HANDLE_BYTES, "handle size");
    DeviceGuard guard(device);
    if (cudaMalloc(dev_ptr, bytes) != cudaSuccess) { cudaGetLastError(); return fail(SGPE_ENOMEM, "device allocation failed"); }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
    if (e != cudaSuccess) {
        cudaFree(*dev_ptr); *dev_ptr = nullptr;
        return fail(SGPE_ECUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
#else
    (void)device;
    *dev_ptr = malloc(bytes);
    if (!*dev_ptr) return fail(SGPE_ENOMEM, "allocation failed");
    memcpy(handle, dev_ptr, sizeof(void*));          // emulation: only meaningful inside one process
#endif
    return 0;
}

int sgpe_ipc_open(int device, const unsigned char handle[SGPE_IPC_HANDLE_BYTES], void** dev_ptr) {
    if (!dev_ptr || !handle) return fail(SGPE_EINVAL, "bad argument");
#ifndef SGPE_EMU
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    SGPE_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
#else
    (void)device;
    memcpy(dev_ptr, handle, sizeof(void*));
#endif
    return 0;
}

int sgpe_ipc_close(void* dev_ptr) {
#ifndef SGPE_EMU
    if (dev_ptr) SGPE_CUDA(cudaIpcCloseMemHandle(dev_ptr));
#else
    (void)dev_ptr;
#endif
    return 0;
}

int sgpe_ipc_free(void* dev_ptr) {
#ifndef SGPE_EMU
    if (dev_ptr) SGPE_CUDA(cudaFree(dev_ptr));
#else
    free(dev_ptr);
#endif
    return 0;
}

int sgpe_plan_create_lines(sgpe_plan** out, int len, int nlines, int n1, int dtype, int device) {
    if (!out) return fail(SGPE_EINVAL, "null output pointer");
    *out = nullptr;
    if (n1 < 1 || len % n1 != 0) return fail(SGPE_EINVAL, "n1 must divide the line length");
    const int n2 = len / n1;
    if (n1 == 1) {
        if (!sgpe::supported_length(len)) return fail(SGPE_EINVAL, "line length must be a power of two in [32, 4096]");
    } else if (!sgpe::supported_length(n1) || n1 > 1024 || !sgpe::supported_length(n2)) {
        return fail(SGPE_EINVAL, "four-step split: n1 in [32, 1024], n2 = len / n1 in [32, 4096], powers of two");
    }
    if (nlines < 1) return fail(SGPE_EINVAL, "nlines must be >= 1");
    if (dtype != SGPE_C128 && dtype != SGPE_C64) return fail(SGPE_EINVAL, "dtype must be 0 (c128) or 1 (c64)");
    DeviceGuard guard(device);
    sgpe_plan* p = new sgpe_plan();
    p->nx = len; p->ny = nlines; p->batch = 1; p->dtype = dtype; p->device = device; p->line_plan = true;
    p->csize = dtype == SGPE_C128 ? 16 : 8;
    p->plane = (long long)len * nlines;
    p->n1 = n1; p->n2 = n2;
    p->max_tiles = 2 * (nlines * n1 > len ? nlines * n1 : len);
    if (p->max_tiles < 1024) p->max_tiles = 1024;
    int rc = 0;
    do {
        if (cudaMalloc((void**)&p->partials, sizeof(double) * 2 * (size_t)p->max_tiles) != cudaSuccess) { rc = SGPE_ENOMEM; break; }
        if (cudaMalloc((void**)&p->counter, sizeof(unsigned) * SGPE_MAX_CHUNKS) != cudaSuccess) { rc = SGPE_ENOMEM; break; }
        if (cudaMalloc((void**)&p->totals, sizeof(double) * 4) != cudaSuccess) { rc = SGPE_ENOMEM; break; }
        if (cudaMalloc((void**)&p->totals_aux, sizeof(double) * 4) != cudaSuccess) { rc = SGPE_ENOMEM; break; }
        cudaMemset(p->counter, 0, sizeof(unsigned) * SGPE_MAX_CHUNKS);
        const int cont = n1 == 1 ? len : n2;
        rc = dtype == SGPE_C128 ? upload_twiddles<double>(&p->tw_x, cont) : upload_twiddles<float>(&p->tw_x, cont);
        if (rc) break;
        if (n1 > 1) {
            rc = dtype == SGPE_C128 ? upload_twiddles<double>(&p->tw_mid, n1) : upload_twiddles<float>(&p->tw_mid, n1);
            if (rc) break;
            rc = dtype == SGPE_C128 ? upload_tw4<double>(p) : upload_tw4<float>(p);
        }
    } while (0);
    if (rc) {
        if (rc == SGPE_ENOMEM) fail(rc, "device allocation failed");
        sgpe_plan_destroy(p);
        return rc;
    }
    *out = p;
    return 0;
}

int sgpe_pass_mid(sgpe_plan* p, void* buf, int pre_tw, int do_inv, int do_pw, double dt_sub, int do_fwd, int post_tw,
                  const double* totals_dev, double global_points, int inner, int scatter, sgpe_stream st) {
    if (!p || !buf) return fail(SGPE_EINVAL, "null argument");
    if (p->n1 <= 1) return fail(SGPE_ESTATE, "not a four-step (long line) plan");
    if (do_pw && (!totals_dev || !(p->grid_set && p->g_set && p->pot_set && p->time_set)))
        return fail(SGPE_ESTATE, "plan not configured for the point-wise operators");
    int rcs = scatter_ready(p, scatter, 2);
    if (rcs) return rcs;
    if (scatter && inner <= 1) return fail(SGPE_EINVAL, "the strided pass stores the exchange only on the column-slab layout");
    DeviceGuard guard(p->device);
    return SGPE_BY_DTYPE(p, run_mid, p, buf, pre_tw != 0, do_inv != 0, do_pw != 0, dt_sub, do_fwd != 0, post_tw != 0,
                         totals_dev, global_points > 0 ? global_points : 1.0, inner, scatter != 0, (cudaStream_t)st);
}

int sgpe_pass_klines(sgpe_plan* p, void* buf, int do_fwd, int has_a, double tau_a, int has_b, double tau_b, int do_inv,
                     double* sums_dev, int scatter, sgpe_stream st) {
    if (!p || !buf) return fail(SGPE_EINVAL, "null argument");
    if ((has_a || has_b) && (!p->kin_set || !p->time_set || !sums_dev)) return fail(SGPE_ESTATE, "kinetic operator / time / sums missing");
    if (p->batch != 1) return fail(SGPE_EINVAL, "line passes are for batch == 1 plans");
    int rcs = scatter_ready(p, scatter, 1);
    if (rcs) return rcs;
    DeviceGuard guard(p->device);
    return SGPE_BY_DTYPE(p, run_klines, p, buf, do_fwd != 0, has_a != 0, tau_a, has_b != 0, tau_b, do_inv != 0,
                         sums_dev, scatter != 0, (cudaStream_t)st);
}

int sgpe_slab_pack(sgpe_plan* p, const void* in, void* out, int lines, int nranks, int chunk, sgpe_stream st) {
    if (!p || !in || !out || in == out) return fail(SGPE_EINVAL, "bad argument");
    DeviceGuard guard(p->device);
    return SGPE_BY_DTYPE(p, run_pack, p, in, out, lines, nranks, chunk, (cudaStream_t)st);
}

int sgpe_slab_unpack(sgpe_plan* p, const void* in, void* out, int nranks, int block_h, int block_w, sgpe_stream st) {
    if (!p || !in || !out || in == out) return fail(SGPE_EINVAL, "bad argument");
    if (block_h % 32 || block_w % 32) return fail(SGPE_EINVAL, "transpose blocks must be multiples of 32");
    DeviceGuard guard(p->device);
    return SGPE_BY_DTYPE(p, run_unpack, p, in, out, nranks, block_h, block_w, (cudaStream_t)st);
}

int sgpe_run_host(sgpe_plan* p, const void* psik_in, void* psik_out, int n_steps, double* pops_host, sgpe_stream st) {
    if (!p || !psik_in || !psik_out) return fail(SGPE_EINVAL, "null argument");
    if (!(p->grid_set && p->g_set && p->kin_set && p->pot_set && p->time_set))
        return fail(SGPE_ESTATE, "set grid, interactions, kinetic, potential and time before stepping");
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)st;
    const size_t bytes = (size_t)p->batch * 2 * p->plane * p->csize;
    const size_t pops_n = (size_t)p->batch * n_steps * 2;
    if (pops_host && pops_n > p->pops_cap) {
        cudaFree(p->pops_buf); p->pops_buf = nullptr; p->pops_cap = 0;
        SGPE_CUDA(cudaMalloc((void**)&p->pops_buf, sizeof(double) * pops_n));
        p->pops_cap = pops_n;
    }
    if (int rc0 = ensure_state(p)) return rc0;
    SGPE_CUDA(cudaMemcpyAsync(p->state, psik_in, bytes, cudaMemcpyHostToDevice, s));
    p->phase = sgpe_plan::KSPACE; p->scale_pending = false; p->pend_slot = -1; p->pend_eslot = -1;
    int rc = sgpe_full_steps(p, n_steps, pops_host ? p->pops_buf : nullptr, 2LL * n_steps, 0, st);
    if (rc) return rc;
    rc = sgpe_store_psik(p, p->state, st);
    if (rc) return rc;
    SGPE_CUDA(cudaMemcpyAsync(psik_out, p->state, bytes, cudaMemcpyDeviceToHost, s));
    if (pops_host && pops_n)
        SGPE_CUDA(cudaMemcpyAsync(pops_host, p->pops_buf, sizeof(double) * pops_n, cudaMemcpyDeviceToHost, s));
    SGPE_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int sgpe_step_accounting(const sgpe_plan* p, uint64_t* algorithmic, uint64_t* actual, int* launches) {
    if (!p) return fail(SGPE_EINVAL, "null plan");
    const uint64_t pts = (uint64_t)p->plane * p->batch;
    if (algorithmic) *algorithmic = pts * (p->dtype == SGPE_C128 ? 768u : 384u);
    if (actual) {
        // per single step: 2 passes x (read + write) x 2 components x csize, plus the operator grids
        uint64_t per_pt = 2 * 2 * 2 * p->csize;
        per_pt += 2 * 8;                                             // kin0, kin1 (column pass)
        per_pt += (p->pot0 == p->pot1) ? 8 : 16;                     // potential (row pass)
        if (p->cpl_mode == SGPE_COUPLING_DENSE) per_pt += 8;
        *actual = 3 * per_pt * pts;
    }
    if (launches) *launches = 6;
    return 0;
}

int sgpe_profile_begin(sgpe_plan* p) {
    if (!p) return fail(SGPE_EINVAL, "null plan");
    p->prof_on = true;
    return 0;
}

int sgpe_profile_end(sgpe_plan* p, double* ms_col, uint64_t* n_col, double* ms_row, uint64_t* n_row) {
    if (!p) return fail(SGPE_EINVAL, "null plan");
    p->prof_on = false;
    double ms[2] = {0.0, 0.0};
    uint64_t cnt[2] = {0, 0};
#ifndef SGPE_EMU
    DeviceGuard guard(p->device);
    for (auto& r : p->prof) {
        SGPE_CUDA(cudaEventSynchronize(r.e1));
        float t = 0.f;
        SGPE_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms[r.kind] += t; cnt[r.kind]++;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    p->prof.clear();
#endif
    if (ms_col) *ms_col = ms[0];
    if (n_col) *n_col = cnt[0];
    if (ms_row) *ms_row = ms[1];
    if (n_row) *n_row = cnt[1];
    return 0;
}

int sgpe_debug_timeline(sgpe_plan* p, unsigned long long* buf_dev) {
    if (!p) return fail(SGPE_EINVAL, "null plan");
    if (p->dbg_kind == 1) { p->dbg_col = buf_dev; if (!buf_dev) p->dbg = nullptr; }
    else { p->dbg = buf_dev; if (!buf_dev) p->dbg_col = nullptr; }
    return 0;
}

int sgpe_launch_count(const sgpe_plan* p, uint64_t* launches) {
    if (!p || !launches) return fail(SGPE_EINVAL, "null argument");
    *launches = p->launches;
    return 0;
}

}  // extern "C"
