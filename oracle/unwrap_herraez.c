/* unwrap_herraez.c — CPU oracle for two-dimensional phase unwrapping.  TEST INFRASTRUCTURE, NOT A PRODUCT PATH
 * (only tests/ may build, load or call it; see oracle/spinor_oracle.py for the rule).
 *
 * What it restates.  The reference unwraps the phase of each real-space component with
 * skimage.restoration.unwrap_phase (spinor_gpe/pspinor/tensor_tools.py:531, reached from
 * TensorPropagator.eng_expect, tensor_propagator.py:304).  scikit-image is a third-party dependency that is not
 * vendored under /root/reference and not installed in this image (pinned scikit-image==0.16.2,
 * requirements.txt:25), so this file restates the PUBLISHED algorithm it implements:
 *
 *   M. A. Herraez, D. R. Burton, M. J. Lalor, M. A. Gdeisat, "Fast two-dimensional phase-unwrapping algorithm
 *   based on sorting by reliability following a noncontinuous path", Appl. Opt. 41, 7437 (2002),
 *
 * in the variant scikit-image ships for 2-D arrays without wrap-around and without a mask (the reference passes a
 * plain ndarray and no keyword arguments):
 *   1. reliability of an interior pixel = H^2 + V^2 + D1^2 + D2^2, the squared second differences of the wrapped
 *      phase along the row, the column and the two diagonals, every first difference wrapped into [-pi, pi];
 *      pixels on the image border get a very large value (scikit-image: 9999999 + rand(), i.e. "least reliable,
 *      in random order"; here 9999999 exactly, ties broken by edge index — every non-corner border pixel then
 *      joins through its interior neighbour whatever the random numbers were, so only the four corners can differ);
 *   2. one edge per horizontally / vertically adjacent pixel pair, reliability = sum of the two pixels', carrying
 *      the integer wrap count between the two pixels;
 *   3. edges sorted by ascending reliability value (most reliable first);
 *   4. pixels are merged into groups edge by edge; joining a group shifts every pixel of the SMALLER group by the
 *      multiple of 2 pi that makes the edge continuous (a lone second pixel joins the first pixel's group, a lone
 *      first pixel joins the second pixel's group; between two proper groups the strictly larger one absorbs);
 *   5. result = wrapped phase + 2 pi * increment.
 *
 * PARITY UNPINNED: without scikit-image there is no golden vector for this step.  The restatement is checked
 * against properties (tests/test_unwrap.py): exact recovery of smooth fields wrapped into (-pi, pi], agreement
 * with numpy.unwrap on fields that vary along one axis, differences that are exact multiples of 2 pi, independence
 * of the result from the merge bookkeeping (this linked-list version vs the union-find of the CUDA library).
 *
 * Data structures follow the paper (explicit pixel groups as linked lists with head / last / next pointers), on
 * purpose different from the product's implementation (offset-carrying union-find, spinor_gpe_b200/csrc).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.141592653589793
#define TWOPI 6.283185307179586
#define BORDER_RELIABILITY 9999999.0

typedef struct pixel_s {
    int increment;              /* multiples of 2 pi to add */
    int group_size;             /* valid at the head of a group */
    double value;               /* wrapped phase */
    double reliability;
    struct pixel_s* head;       /* first pixel of the group this pixel belongs to */
    struct pixel_s* last;       /* valid at the head: last pixel of the group */
    struct pixel_s* next;       /* next pixel of the group, NULL at the end */
} pixel_t;

typedef struct {
    double reliability;
    pixel_t* first;
    pixel_t* second;
    int increment;              /* wrap count between first and second */
    int64_t index;              /* creation order, the tie-break of the sort */
} edge_t;

static double wrap(double d) {
    if (d > PI) return d - TWOPI;
    if (d < -PI) return d + TWOPI;
    return d;
}

static int find_wrap(double left, double right) {
    const double d = left - right;
    if (d > PI) return -1;
    if (d < -PI) return 1;
    return 0;
}

static int edge_cmp(const void* a, const void* b) {
    const edge_t* x = (const edge_t*)a;
    const edge_t* y = (const edge_t*)b;
    if (x->reliability < y->reliability) return -1;
    if (x->reliability > y->reliability) return 1;
    return (x->index > y->index) - (x->index < y->index);
}

/* wrapped, unwrapped: [height][width] row-major doubles.  increments (nullable): the integer field.
 * Returns 0, or -1 when out of memory. */
int unwrap2d_oracle(const double* wrapped, double* unwrapped, int32_t* increments, int width, int height) {
    const int64_t n = (int64_t)width * height;
    const int64_t n_edges = (int64_t)(width - 1) * height + (int64_t)width * (height - 1);
    pixel_t* px = (pixel_t*)malloc(sizeof(pixel_t) * (size_t)n);
    edge_t* ed = (edge_t*)malloc(sizeof(edge_t) * (size_t)(n_edges > 0 ? n_edges : 1));
    if (!px || !ed) { free(px); free(ed); return -1; }

    for (int64_t i = 0; i < n; i++) {
        px[i].increment = 0; px[i].group_size = 1; px[i].value = wrapped[i];
        px[i].reliability = BORDER_RELIABILITY;
        px[i].head = &px[i]; px[i].last = &px[i]; px[i].next = NULL;
    }
    /* 1. reliabilities of the interior */
    for (int i = 1; i < height - 1; i++) {
        for (int j = 1; j < width - 1; j++) {
            const double* w = wrapped + (int64_t)i * width + j;
            const double h = wrap(w[-1] - w[0]) - wrap(w[0] - w[1]);
            const double v = wrap(w[-width] - w[0]) - wrap(w[0] - w[width]);
            const double d1 = wrap(w[-width - 1] - w[0]) - wrap(w[0] - w[width + 1]);
            const double d2 = wrap(w[-width + 1] - w[0]) - wrap(w[0] - w[width - 1]);
            px[(int64_t)i * width + j].reliability = h * h + v * v + d1 * d1 + d2 * d2;
        }
    }
    /* 2. edges: all horizontal ones row by row, then all vertical ones */
    int64_t e = 0;
    for (int i = 0; i < height; i++)
        for (int j = 0; j < width - 1; j++) {
            pixel_t* a = &px[(int64_t)i * width + j];
            ed[e].first = a; ed[e].second = a + 1;
            ed[e].reliability = a->reliability + (a + 1)->reliability;
            ed[e].increment = find_wrap(a->value, (a + 1)->value);
            ed[e].index = e; e++;
        }
    for (int i = 0; i < height - 1; i++)
        for (int j = 0; j < width; j++) {
            pixel_t* a = &px[(int64_t)i * width + j];
            ed[e].first = a; ed[e].second = a + width;
            ed[e].reliability = a->reliability + (a + width)->reliability;
            ed[e].increment = find_wrap(a->value, (a + width)->value);
            ed[e].index = e; e++;
        }
    /* 3. most reliable (smallest value) first */
    qsort(ed, (size_t)n_edges, sizeof(edge_t), edge_cmp);
    /* 4. gather the pixels into groups */
    for (int64_t k = 0; k < n_edges; k++) {
        pixel_t* p1 = ed[k].first;
        pixel_t* p2 = ed[k].second;
        if (p1->head == p2->head) continue;
        if (p2->next == NULL && p2->head == p2) {                   /* p2 is alone: it joins p1's group */
            p1->head->last->next = p2;
            p1->head->last = p2;
            p1->head->group_size++;
            p2->head = p1->head;
            p2->increment = p1->increment - ed[k].increment;
        } else if (p1->next == NULL && p1->head == p1) {            /* p1 is alone: it joins p2's group */
            p2->head->last->next = p1;
            p2->head->last = p1;
            p2->head->group_size++;
            p1->head = p2->head;
            p1->increment = p2->increment + ed[k].increment;
        } else {
            pixel_t* g1 = p1->head;
            pixel_t* g2 = p2->head;
            if (g1->group_size > g2->group_size) {                  /* group 2 joins group 1 */
                const int shift = p1->increment - ed[k].increment - p2->increment;
                g1->last->next = g2;
                g1->last = g2->last;
                g1->group_size += g2->group_size;
                for (pixel_t* q = g2; q != NULL; q = q->next) { q->head = g1; q->increment += shift; }
            } else {                                                /* group 1 joins group 2 */
                const int shift = p2->increment + ed[k].increment - p1->increment;
                g2->last->next = g1;
                g2->last = g1->last;
                g2->group_size += g1->group_size;
                for (pixel_t* q = g1; q != NULL; q = q->next) { q->head = g2; q->increment += shift; }
            }
        }
    }
    /* 5. unwrap */
    for (int64_t i = 0; i < n; i++) {
        unwrapped[i] = px[i].value + TWOPI * (double)px[i].increment;
        if (increments) increments[i] = px[i].increment;
    }
    free(px); free(ed);
    return 0;
}
