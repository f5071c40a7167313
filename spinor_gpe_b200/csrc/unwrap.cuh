// unwrap.cuh — device side of the two-dimensional phase unwrapping the energy expectation needs.
//
// The reference unwraps the phase of each real-space component with skimage.restoration.unwrap_phase
// (tensor_tools.py:531, from TensorPropagator.eng_expect, tensor_propagator.py:304): the algorithm of Herraez,
// Burton, Lalor and Gdeisat, Appl. Opt. 41, 7437 (2002) — "sorting by reliability following a non-continuous path".
// Its data-parallel parts run here, one thread per pixel:
//   unwrap_angle_pass  : phi = atan2(Im psi, Re psi)                                     (np.angle, :529)
//   unwrap_reliab_pass : reliability = H^2 + V^2 + D1^2 + D2^2 of the wrapped second differences; border pixels get
//                        9999999 ("least reliable"; scikit-image adds rand() there, which only decides how the four
//                        corner pixels attach — every other border pixel joins through its interior neighbour)
//   unwrap_edge_pass   : one sort key (the reliability sum of the two pixels, as order-preserving integer bits) and
//                        one payload (edge id and wrap count) per horizontal / vertical pixel pair
//   (radix sort of the edges by key: cub, in sgpe_api.cu)
//   unwrap_apply_pass  : phi + 2 pi * increment, optionally zeroed where the density is below 1e-6 of its maximum
//                        (tensor_tools.py:538)
// The region merging between sort and apply is Kruskal's algorithm on the reliability-sorted edges.  Its RESULT splits
// into two parts of very different nature:
//   * the minimum spanning tree of the pixel grid under the (unique) edge ranks and, along its edges, the relative
//     multiples of 2 pi of every pixel pair — independent of the order in which the tree is grown, so it is built on
//     the device by Boruvka rounds over an offset-carrying union-find (unwrap_minedge_pass / unwrap_hook_pass /
//     unwrap_adopt_pass / unwrap_compress_pass below; the tree is the same one Kruskal finds because ranks are unique);
//   * the ONE pixel group that never moves (the published merge rules let the larger group keep its values: the global
//     2 pi offset of the field, visible at the mask edge of the energy) — this depends on the group sizes at every merge
//     in rank order.  Taken literally that is a sequential pass over the N - 1 tree edges (unwrap_anchor in
//     sgpe_api.cu: sizes only, no offset bookkeeping; used when there are many planes, one per host core); on the
//     device it is found level by level — the first merge that creates a group of more than half the pixels decides,
//     recursively inside the winning half — by a bisection over the rank threshold with a size-carrying lock-free
//     union-find (unwrap_level_*_pass at the end of this file).
// Option "unwrap_merge" = 1 keeps the whole merging on the host (offset-carrying union-find over all edges), for
// cross-checks.
//
// The arithmetic that decides the ORDER of the edges is written with explicit round-to-nearest multiplies and adds
// (no FMA contraction) so that keys are bit-identical to a plain C evaluation of the same expressions.
#pragma once

#include "fft_core.cuh"

namespace sgpe {

constexpr double kUnwrapPi = 3.141592653589793;
constexpr double kUnwrapTwoPi = 6.283185307179586;
constexpr double kUnwrapBorder = 9999999.0;

SGPE_DI double unwrap_wrap(double d) {
    if (d > kUnwrapPi) return d - kUnwrapTwoPi;
    if (d < -kUnwrapPi) return d + kUnwrapTwoPi;
    return d;
}
// multiples of 2 pi the right / lower pixel of an edge needs relative to the left / upper one, negated
SGPE_DI int unwrap_find_wrap(double left, double right) {
    const double d = left - right;
    if (d > kUnwrapPi) return -1;
    if (d < -kUnwrapPi) return 1;
    return 0;
}
SGPE_DI double unwrap_second_diff_sq(double before, double centre, double after) {
    const double s = unwrap_wrap(before - centre) - unwrap_wrap(centre - after);
    return __dmul_rn(s, s);
}

template <typename T>
__global__ void __launch_bounds__(256) unwrap_angle_pass(const typename cx_of<T>::type* psi, long long total,
                                                         double* phi) {
    typedef typename cx_of<T>::type C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const C z = psi[i];
        phi[i] = atan2((double)z.y, (double)z.x);
    }
}

// per-plane maximum of |psi|^2 as the bits of a non-negative double (integer order == floating-point order);
// grid (blocks, planes), maxbits zeroed by the caller
template <typename T>
__global__ void __launch_bounds__(256) unwrap_maxdens_pass(const typename cx_of<T>::type* psi, long long plane,
                                                           unsigned long long* maxbits) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);
    const C* p = psi + (long long)blockIdx.y * plane;
    double mx = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane;
         i += (long long)gridDim.x * blockDim.x) {
        const C z = p[i];
        const double d = (double)z.x * z.x + (double)z.y * z.y;
        mx = d > mx ? d : mx;
    }
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = red[threadIdx.x] > red[threadIdx.x + s] ? red[threadIdx.x] : red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(&maxbits[blockIdx.y], (unsigned long long)__double_as_longlong(red[0]));
}

__global__ void __launch_bounds__(256) unwrap_reliab_pass(const double* phi, int nx, int ny, double* rel) {
    const long long plane = (long long)nx * ny;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < plane;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / nx), j = (int)(idx - (long long)i * nx);
        double r = kUnwrapBorder;
        if (i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
            const double* w = phi + idx;
            const double c = w[0];
            const double h = unwrap_second_diff_sq(w[-1], c, w[1]);
            const double v = unwrap_second_diff_sq(w[-nx], c, w[nx]);
            const double d1 = unwrap_second_diff_sq(w[-nx - 1], c, w[nx + 1]);
            const double d2 = unwrap_second_diff_sq(w[-nx + 1], c, w[nx - 1]);
            r = __dadd_rn(__dadd_rn(__dadd_rn(h, v), d1), d2);
        }
        rel[idx] = r;
    }
}

// Edge e < ny (nx - 1): the horizontal pair (i, j)-(i, j + 1), e = i (nx - 1) + j; then the vertical pairs
// (i, j)-(i + 1, j), e = ny (nx - 1) + i nx + j.  Payload = e << 2 | (wrap count + 1).
__global__ void __launch_bounds__(256) unwrap_edge_pass(const double* phi, const double* rel, int nx, int ny,
                                                        unsigned long long* keys, unsigned* vals) {
    const long long plane = (long long)nx * ny;
    const long long n_horizontal = (long long)ny * (nx - 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < plane;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / nx), j = (int)(idx - (long long)i * nx);
        const double c = phi[idx], rc = rel[idx];
        if (j < nx - 1) {
            const long long e = (long long)i * (nx - 1) + j;
            keys[e] = (unsigned long long)__double_as_longlong(__dadd_rn(rc, rel[idx + 1]));
            vals[e] = ((unsigned)e << 2) | (unsigned)(unwrap_find_wrap(c, phi[idx + 1]) + 1);
        }
        if (i < ny - 1) {
            const long long e = n_horizontal + idx;
            keys[e] = (unsigned long long)__double_as_longlong(__dadd_rn(rc, rel[idx + nx]));
            vals[e] = ((unsigned)e << 2) | (unsigned)(unwrap_find_wrap(c, phi[idx + nx]) + 1);
        }
    }
}

// out = phi + 2 pi inc; with psi and maxbits given, 0 where |psi|^2 < 1e-6 max|psi|^2 of the plane
template <typename T>
__global__ void __launch_bounds__(256) unwrap_apply_pass(const typename cx_of<T>::type* psi, const double* phi,
                                                         const int* inc, const unsigned long long* maxbits,
                                                         long long plane, long long total, double* out) {
    typedef typename cx_of<T>::type C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        double v = __dadd_rn(phi[i], __dmul_rn(kUnwrapTwoPi, (double)inc[i]));
        if (psi != nullptr) {
            const C z = psi[i];
            const double n = (double)z.x * z.x + (double)z.y * z.y;
            const double thr = __longlong_as_double((long long)maxbits[i / plane]) * 1e-6;
            if (n < thr) v = 0.0;
        }
        out[i] = v;
    }
}

// ---- region merging on the device: Boruvka rounds over an offset-carrying union-find --------------------------------
// node[v] = parent(v) in the low word, increment(v) - increment(parent(v)) in the high word (one 64-bit word, so a
// reader always sees a consistent pair).  Between rounds every pixel points at the root of its group (a star).
constexpr unsigned kUnwrapNoEdge = 0xffffffffu;
constexpr unsigned kUnwrapTreeBit = 0x80000000u;     // set in the sorted payload of an edge once it joins the tree

SGPE_DI unsigned long long unwrap_pack(unsigned parent, int off) {
    return ((unsigned long long)(unsigned)off << 32) | parent;
}
SGPE_DI unsigned unwrap_parent(unsigned long long n) { return (unsigned)(n & 0xffffffffull); }
SGPE_DI int unwrap_off(unsigned long long n) { return (int)(unsigned)(n >> 32); }
// the two pixels of edge e (numbering of unwrap_edge_pass)
SGPE_DI void unwrap_edge_ends(unsigned e, int nx, unsigned n_horizontal, unsigned* p1, unsigned* p2) {
    if (e < n_horizontal) {
        const unsigned i = e / (unsigned)(nx - 1), j = e - i * (unsigned)(nx - 1);
        *p1 = i * (unsigned)nx + j; *p2 = *p1 + 1u;
    } else {
        *p1 = e - n_horizontal; *p2 = *p1 + (unsigned)nx;
    }
}

// rank_of[e] = position of edge e in the sorted order (ranks are unique: ties were broken by edge id in the sort)
__global__ void __launch_bounds__(256) unwrap_rank_pass(const unsigned* sorted, long long n_edges, unsigned* rank_of) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_edges;
         k += (long long)gridDim.x * blockDim.x)
        rank_of[(sorted[k] & ~kUnwrapTreeBit) >> 2] = (unsigned)k;
}

__global__ void __launch_bounds__(256) unwrap_forest_init_pass(long long plane, unsigned long long* node, unsigned* best) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < plane;
         v += (long long)gridDim.x * blockDim.x) {
        node[v] = unwrap_pack((unsigned)v, 0);
        best[v] = kUnwrapNoEdge;
    }
}

// every pixel offers the lowest-ranked of its (up to four) edges that leave its group to the group's root
__global__ void __launch_bounds__(256) unwrap_minedge_pass(const unsigned long long* node, const unsigned* rank_of,
                                                           int nx, int ny, unsigned* best) {
    const long long plane = (long long)nx * ny;
    const unsigned n_horizontal = (unsigned)ny * (unsigned)(nx - 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < plane;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / nx), j = (int)(idx - (long long)i * nx);
        const unsigned root = unwrap_parent(node[idx]);
        unsigned m = kUnwrapNoEdge;
        if (j < nx - 1 && unwrap_parent(node[idx + 1]) != root) { const unsigned r = rank_of[(unsigned)i * (unsigned)(nx - 1) + j]; m = r < m ? r : m; }
        if (j > 0 && unwrap_parent(node[idx - 1]) != root) { const unsigned r = rank_of[(unsigned)i * (unsigned)(nx - 1) + j - 1]; m = r < m ? r : m; }
        if (i < ny - 1 && unwrap_parent(node[idx + nx]) != root) { const unsigned r = rank_of[n_horizontal + (unsigned)idx]; m = r < m ? r : m; }
        if (i > 0 && unwrap_parent(node[idx - nx]) != root) { const unsigned r = rank_of[n_horizontal + (unsigned)(idx - nx)]; m = r < m ? r : m; }
        if (m != kUnwrapNoEdge && m < best[root]) atomicMin(&best[root], m);
    }
}

// every root with an outgoing edge hangs itself below the root on the other side of its lowest-ranked one (written to
// pend[], adopted by the next pass: node[] stays a forest of stars while this pass reads it).  Two groups that chose
// the same edge: the root with the smaller index stays.  Ranks are unique, so there is no longer cycle.
__global__ void __launch_bounds__(256) unwrap_hook_pass(const unsigned long long* node, const unsigned* best,
                                                        unsigned* sorted, int nx, int ny, unsigned long long* pend,
                                                        unsigned* hooked) {
    const long long plane = (long long)nx * ny;
    const unsigned n_horizontal = (unsigned)ny * (unsigned)(nx - 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < plane;
         idx += (long long)gridDim.x * blockDim.x) {
        const unsigned v = (unsigned)idx;
        if (unwrap_parent(node[idx]) != v) continue;
        unsigned long long out = unwrap_pack(v, 0);
        const unsigned k = best[idx];
        if (k != kUnwrapNoEdge) {
            const unsigned payload = sorted[k] & ~kUnwrapTreeBit;
            const int wraps = (int)(payload & 3u) - 1;            // increment(p1) - increment(p2)
            unsigned p1, p2;
            unwrap_edge_ends(payload >> 2, nx, n_horizontal, &p1, &p2);
            const unsigned long long n1 = node[p1], n2 = node[p2];
            const bool first_is_mine = unwrap_parent(n1) == v;
            const unsigned other_root = first_is_mine ? unwrap_parent(n2) : unwrap_parent(n1);
            const int a_mine = first_is_mine ? unwrap_off(n1) : unwrap_off(n2);
            const int a_other = first_is_mine ? unwrap_off(n2) : unwrap_off(n1);
            if (!(best[other_root] == k && v < other_root)) {
                // increment(v) - increment(other_root)
                out = unwrap_pack(other_root, (first_is_mine ? wraps : -wraps) + a_other - a_mine);
                atomicOr(&sorted[k], kUnwrapTreeBit);
                atomicAdd(hooked, 1u);
            }
        }
        pend[idx] = out;
    }
}

__global__ void __launch_bounds__(256) unwrap_adopt_pass(long long plane, const unsigned long long* pend,
                                                         unsigned long long* node, unsigned* best) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < plane;
         v += (long long)gridDim.x * blockDim.x) {
        if (unwrap_parent(node[v]) == (unsigned)v) node[v] = pend[v];
        best[v] = kUnwrapNoEdge;
    }
}

// pointer jumping until every pixel points at a root again.  Each thread rewrites only its own word, and every word is
// at all times a true statement "increment(v) - increment(parent) = off" about SOME ancestor, so concurrent readers
// may see any mixture of old and new words.
__global__ void __launch_bounds__(256) unwrap_compress_pass(long long plane, unsigned long long* node) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < plane;
         v += (long long)gridDim.x * blockDim.x) {
        unsigned long long mine = node[v];
        for (;;) {
            const unsigned p = unwrap_parent(mine);
            const unsigned long long up = __ldcg(&node[p]);
            if (unwrap_parent(up) == p) break;
            mine = unwrap_pack(unwrap_parent(up), unwrap_off(mine) + unwrap_off(up));
            node[v] = mine;
        }
    }
}

// raw[v] = increment(v) - increment(root of the tree)
__global__ void __launch_bounds__(256) unwrap_offsets_pass(long long plane, const unsigned long long* node, int* raw) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < plane;
         v += (long long)gridDim.x * blockDim.x)
        raw[v] = unwrap_off(node[v]);
}

// the tree edges in rank order for the host's anchor pass: first pixel, bit 31 = vertical edge; others kUnwrapNoEdge
__global__ void __launch_bounds__(256) unwrap_tree_edges_pass(const unsigned* sorted, long long n_edges, int nx, int ny,
                                                              unsigned* cand) {
    const unsigned n_horizontal = (unsigned)ny * (unsigned)(nx - 1);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_edges;
         k += (long long)gridDim.x * blockDim.x) {
        const unsigned s = sorted[k];
        unsigned out = kUnwrapNoEdge;
        if (s & kUnwrapTreeBit) {
            unsigned p1, p2;
            const unsigned e = (s & ~kUnwrapTreeBit) >> 2;
            unwrap_edge_ends(e, nx, n_horizontal, &p1, &p2);
            out = p1 | (e >= n_horizontal ? kUnwrapTreeBit : 0u);
        }
        cand[k] = out;
    }
}

// inc[v] -= inc[anchor]: the anchor pixel's group is the one that never moved.  The anchor's own entry is left alone
// here (every thread of the grid reads it) and zeroed by unwrap_anchor_zero_pass afterwards.
__global__ void __launch_bounds__(256) unwrap_anchor_pass(long long plane, long long anchor, int* inc) {
    const int base = inc[anchor];
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < plane;
         v += (long long)gridDim.x * blockDim.x)
        if (v != anchor) inc[v] -= base;
}
__global__ void unwrap_anchor_zero_pass(long long anchor, int* inc) { if (threadIdx.x == 0) inc[anchor] = 0; }

// ---- the pixel group that never moves, found on the device ----------------------------------------------------------
// Kruskal keeps the values of the LARGER group at every merge, so inside any group X that Kruskal ever formed (a
// connected piece of the spanning tree whose inner edges all rank below its outer ones) the surviving lineage is settled
// by one merge: the first one that creates a group of more than |X| / 2 pixels.  From then on that group outweighs
// everything else in X and wins every merge; before, it did not exist.  Its two halves A and B (each at most |X| / 2)
// are again groups Kruskal formed, and the published rule for (|A|, |B|) names the half whose values survive: recurse
// into it.  The size at least halves per level, and "does a group of more than |X| / 2 pixels exist once the edges up
// to rank T are in" is monotone in T, so each level is a bisection over T with a lock-free union-find over the tree
// edges of X (ECL-CC style hooking by index) that carries the group sizes along.  The bisection keeps the forest of its
// lower bound and only adds the edges between the bounds.  Small X go to the host (a few thousand edges).
struct UnwrapEdge { unsigned k, t; };      // position in the rank-sorted tree-edge list; first pixel | vertical << 31

SGPE_DI unsigned unwrap_uf_find(unsigned* parent, unsigned x) {
    unsigned curr = __ldcg(&parent[x]);
    if (curr != x) {
        unsigned prev = x, next;
        while (curr != (next = __ldcg(&parent[curr]))) { parent[prev] = next; prev = curr; curr = next; }
    }
    return curr;
}
// Lock-free union (the larger root index hangs below the smaller, ECL-CC style) with group sizes kept at the roots.
// A root that has just been hung below another hands over what it
// holds with an atomic exchange; whoever adds to a node afterwards looks again whether that node is still a root and,
// if not, takes back what sits there and passes it on.  Either the absorbing thread's exchange comes after an addition
// (and carries it along) or the adder sees the node absorbed (and forwards it itself): pixels are never lost or counted
// twice, and when the pass has ended every size sits at a root.  An addition that lands on a root and lifts it above
// `report_above` reports the new total (the last one to land on a group reports its final size).
SGPE_DI void unwrap_uf_union_sized(unsigned* parent, unsigned* size, unsigned a, unsigned b, unsigned* largest,
                                   unsigned report_above) {
    a = unwrap_uf_find(parent, a);
    b = unwrap_uf_find(parent, b);
    while (a != b) {
        // who hangs below whom is decided by a scrambled index (a bijection of 32-bit integers: still a strict total
        // order, so no cycle can form): pixel indices follow the rows of the image, and on smooth fields so do the
        // groups — linking by the plain index grows long chains there, linking by the scrambled one behaves like
        // random linking
        if (a * 2654435761u < b * 2654435761u) { const unsigned t = a; a = b; b = t; }
        const unsigned seen = atomicCAS(&parent[a], a, b);
        if (seen == a) {
            unsigned carry = atomicExch(&size[a], 0u);
            unsigned to = b;
            while (carry) {
                const unsigned now = atomicAdd(&size[to], carry) + carry;
                __threadfence();                                  // the addition is out before the look at the parent
                const unsigned up = __ldcg(&parent[to]);
                if (up == to) { if (now > report_above) atomicMax(largest, now); break; }
                carry = atomicExch(&size[to], 0u);
                to = up;
            }
            break;
        }
        a = seen;
    }
}

// The bisection of a level runs without the host: ctrl = {lo, hi, which forest holds the state at lo} lives on the
// device, the passes of a probe read it (and do nothing once hi - lo <= 1), unwrap_level_step_pass
// moves a bound after each probe.  The host enqueues ceil(log2(hi - lo)) probes and reads the outcome once per level.
struct UnwrapProbe { long long lo, mid; unsigned* snap; unsigned* work; unsigned* snap_size; unsigned* work_size; bool live; };
SGPE_DI UnwrapProbe unwrap_probe(const long long* ctrl, unsigned* forest_a, unsigned* forest_b, unsigned* size_a, unsigned* size_b) {
    UnwrapProbe q;
    const long long hi = ctrl[1];
    q.lo = ctrl[0];
    q.live = hi - q.lo > 1;
    q.mid = q.lo + (hi - q.lo) / 2;
    const bool flipped = ctrl[2] != 0;
    q.snap = flipped ? forest_b : forest_a;
    q.work = flipped ? forest_a : forest_b;
    q.snap_size = flipped ? size_b : size_a;
    q.work_size = flipped ? size_a : size_b;
    return q;
}
// level start: no edge in (lo = -1: every pixel its own group of size 1, in forest_a / size_a)
__global__ void unwrap_level_begin_pass(long long lo, long long hi, long long* ctrl, unsigned* result) {
    if (threadIdx.x == 0) { ctrl[0] = lo; ctrl[1] = hi; ctrl[2] = 0; ctrl[3] = 0; }
    if (threadIdx.x < 8) result[threadIdx.x] = 0u;
}
// after a probe: result[0] = size of the group of more than nv / 2 pixels with the edges up to mid in, or 0
__global__ void unwrap_level_step_pass(long long nv, long long* ctrl, unsigned* result) {
    if (threadIdx.x != 0) return;
    const long long lo = ctrl[0], hi = ctrl[1];
    if (hi - lo > 1) {
        const long long mid = lo + (hi - lo) / 2;
        if (2ll * (long long)result[0] > nv) ctrl[1] = mid;
        else { ctrl[0] = mid; ctrl[2] ^= 1; }
    }
    result[0] = 0;
}

__global__ void __launch_bounds__(256) unwrap_level0_pass(const unsigned* tree, long long plane, unsigned* vl, UnwrapEdge* el) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < plane;
         v += (long long)gridDim.x * blockDim.x) {
        vl[v] = (unsigned)v;
        if (v + 1 < plane) { UnwrapEdge e; e.k = (unsigned)v; e.t = tree[v]; el[v] = e; }
    }
}
// ctrl == nullptr: forest_a[v] = v, size_a[v] = 1.  In a probe: work forest and sizes = copy of those at lo.
__global__ void __launch_bounds__(256) unwrap_level_reset_pass(const unsigned* vl, long long nv, const long long* ctrl,
                                                               unsigned* forest_a, unsigned* forest_b, unsigned* size_a,
                                                               unsigned* size_b) {
    if (!ctrl) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
            const unsigned v = vl[i];
            forest_a[v] = v;
            size_a[v] = 1u;
        }
        return;
    }
    const UnwrapProbe q = unwrap_probe(ctrl, forest_a, forest_b, size_a, size_b);
    if (!q.live) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
        const unsigned v = vl[i];
        q.work[v] = q.snap[v];
        q.work_size[v] = q.snap_size[v];
    }
}
// a probe: joins the ends of the level's edges with lo < k <= mid in the work forest, sizes carried along;
// result[0] (preset to 0) becomes the size of the group of more than nv / 2 pixels, if the probe creates one
__global__ void __launch_bounds__(256) unwrap_level_union_pass(const UnwrapEdge* el, long long ne, const long long* ctrl,
                                                               int nx, unsigned* forest_a, unsigned* forest_b,
                                                               unsigned* size_a, unsigned* size_b, long long nv,
                                                               unsigned* result) {
    const UnwrapProbe q = unwrap_probe(ctrl, forest_a, forest_b, size_a, size_b);
    if (!q.live) return;
    const unsigned half = (unsigned)(nv / 2);                // 2 * size > nv  <=>  size > floor(nv / 2)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += (long long)gridDim.x * blockDim.x) {
        const UnwrapEdge e = el[i];
        if ((long long)e.k <= q.lo || (long long)e.k > q.mid) continue;
        const unsigned p1 = e.t & ~kUnwrapTreeBit;
        unwrap_uf_union_sized(q.work, q.work_size, p1, p1 + ((e.t & kUnwrapTreeBit) ? (unsigned)nx : 1u), result, half);
    }
}
// the two groups edge T = hi joins, in the forest at lo = T - 1: result[1..4] = root and size on the first pixel's
// side, root and size on the second's
__global__ void __launch_bounds__(256) unwrap_level_sides_pass(const UnwrapEdge* el, long long ne, const long long* ctrl, int nx,
                                                               unsigned* forest_a, unsigned* forest_b, unsigned* size_a,
                                                               unsigned* size_b, unsigned* result) {
    const UnwrapProbe q = unwrap_probe(ctrl, forest_a, forest_b, size_a, size_b);
    const long long T = ctrl[1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += (long long)gridDim.x * blockDim.x) {
        const UnwrapEdge e = el[i];
        if ((long long)e.k != T) continue;
        const unsigned p1 = e.t & ~kUnwrapTreeBit;
        const unsigned a = unwrap_uf_find(q.snap, p1);
        const unsigned b = unwrap_uf_find(q.snap, p1 + ((e.t & kUnwrapTreeBit) ? (unsigned)nx : 1u));
        result[1] = a; result[2] = q.snap_size[a]; result[3] = b; result[4] = q.snap_size[b];
    }
}
// next level: the pixels of group `keep` and the edges below rank T inside it (appended in any order;
// result[5] / result[6] count them)
SGPE_DI unsigned unwrap_append_slot(bool take, unsigned* counter) {
#ifdef SGPE_EMU
    return take ? atomicAdd(counter, 1u) : 0u;
#else
    const unsigned votes = __ballot_sync(0xffffffffu, take);
    if (!take) return 0u;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned leader = (unsigned)(__ffs(votes) - 1);
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(votes));
    base = __shfl_sync(votes, base, (int)leader);
    return base + (unsigned)__popc(votes & ((1u << lane) - 1u));
#endif
}
__global__ void __launch_bounds__(256) unwrap_level_select_pass(const unsigned* vl, long long nv, const UnwrapEdge* el,
                                                                long long ne, unsigned T, unsigned keep, unsigned* parent,
                                                                unsigned* vl_out, UnwrapEdge* el_out, unsigned* result) {
    const long long span = nv > ne ? nv : ne;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < span; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        bool take = false;
        unsigned v = 0;
        if (i < nv) { v = vl[i]; take = unwrap_uf_find(parent, v) == keep; }
        unsigned slot = unwrap_append_slot(take, &result[5]);
        if (take) vl_out[slot] = v;
        take = false;
        UnwrapEdge e; e.k = 0; e.t = 0;
        if (i < ne) { e = el[i]; take = e.k < T && unwrap_uf_find(parent, e.t & ~kUnwrapTreeBit) == keep; }
        slot = unwrap_append_slot(take, &result[6]);
        if (take) el_out[slot] = e;
    }
}

}  // namespace sgpe
