// Anchor handle + anchor-generator kernel (K0).
// Reference: nets/ron_vgg_320.py:285-355 (RON rule), nets/ssd_vgg_512.py:286-358 (SSD rule),
// nets/ssd_common.py:371-402 (re-derived encode anchors, per-anchor borders) and :103-115
// (second-trip corners, inside mask).  SURVEY.md Appendix A.1-A.3.
#include <math.h>
#include <math_constants.h>
#include <string.h>

#include <algorithm>

#include "encode_common.cuh"

namespace ronk {

struct AnchorGenParams {
    LayerTable tab;
    float step[kMaxLayers];
    float lo_y[kMaxLayers], lo_x[kMaxLayers], hi_y[kMaxLayers], hi_x[kMaxLayers];
    float h[kMaxShapes], w[kMaxShapes];
    float img_h, img_w, offset;
    int has_border;
};

// One thread per anchor.  The centre grid is three float32 ops in the reference's order
// ((i + offset) * step) / img; per-shape h, w arrive already rounded to float32 (the
// reference evaluates them in Python double and stores to a float32 array).
__global__ void __launch_bounds__(256)
anchor_gen_kernel(const __grid_constant__ AnchorGenParams p, float4* __restrict__ dec,
                  float4* __restrict__ enc, float4* __restrict__ cor, uint8_t* __restrict__ inside) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.tab.N) return;
    int l = layer_of(p.tab, n);
    int local = n - p.tab.offs[l];
    int A = p.tab.A[l];
    int a = local % A;
    int cell = local / A;
    int j = cell % p.tab.W[l];
    int i = cell / p.tab.W[l];
    float y = (((float)i + p.offset) * p.step[l]) / p.img_h;
    float x = (((float)j + p.offset) * p.step[l]) / p.img_w;
    float h = p.h[p.tab.hw_off[l] + a];
    float w = p.w[p.tab.hw_off[l] + a];
    dec[n] = make_float4(y, x, h, w);
    // first trip: corners from the original anchors (ssd_common.py:375-378)
    float ymin_ = y - h / 2.f, xmin_ = x - w / 2.f, ymax_ = y + h / 2.f, xmax_ = x + w / 2.f;
    // re-derived centre/size (ssd_common.py:381)
    float cy = (ymin_ + ymax_) / 2.f, cx = (xmin_ + xmax_) / 2.f;
    float hh = ymax_ - ymin_, ww = xmax_ - xmin_;
    enc[n] = make_float4(cy, cx, hh, ww);
    // second trip (ssd_common.py:105-108)
    float ymin = cy - hh / 2.f, xmin = cx - ww / 2.f, ymax = cy + hh / 2.f, xmax = cx + ww / 2.f;
    cor[n] = make_float4(ymin, xmin, ymax, xmax);
    bool in = true;
    if (p.has_border)
        in = (ymin >= p.lo_y[l]) && (xmin >= p.lo_x[l]) && (ymax < p.hi_y[l]) && (xmax < p.hi_x[l]);
    inside[n] = in ? 1 : 0;
}

// Arbitrary flattened anchors: corners + inside mask from given (y, x, h, w) and per-anchor borders.
__global__ void __launch_bounds__(256)
anchor_flat_kernel(const float4* __restrict__ yxhw, const int* __restrict__ border, int N, int img_h, int img_w,
                   float4* __restrict__ cor, uint8_t* __restrict__ inside) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float4 a = yxhw[n];
    float ymin = a.x - a.z / 2.f, xmin = a.y - a.w / 2.f, ymax = a.x + a.z / 2.f, xmax = a.y + a.w / 2.f;
    cor[n] = make_float4(ymin, xmin, ymax, xmax);
    bool in = true;
    if (border) {
        double b = (double)border[n];
        float lo_y = (float)(-b * 1. / img_h), lo_x = (float)(-b * 1. / img_w);
        float hi_y = (float)((img_h + b) * 1. / img_h), hi_x = (float)((img_w + b) * 1. / img_w);
        in = (ymin >= lo_y) && (xmin >= lo_x) && (ymax < hi_y) && (xmax < hi_x);
    }
    inside[n] = in ? 1 : 0;
}

// Row / column tables, inside rectangles and work items of the grid kernel (match_encode_grid.cu).  Every property the
// kernel relies on is CHECKED here against the per-anchor tables the generator kernel produced: corners separable bit for
// bit, both corner sequences monotone, inside mask = rows x columns, both contiguous.  Any failure leaves grid_ok = 0.
static cudaError_t build_grid(ronk_anchors* h, const std::vector<float>& cor, const std::vector<uint8_t>& in,
                              const std::vector<int>& idx) {
    h->grid_ok = 0;
    if (h->kind < 0) return cudaSuccess;
    const LayerTable& t = h->tab;
    auto bits = [](float v) { uint32_t u; memcpy(&u, &v, 4); return u; };
    std::vector<float> rowtab, coltab;
    std::vector<int> planes;
    std::vector<int> rt_off(t.L), ct_off(t.L), pl_off(t.L), max_cells(t.L, 0), n_in(t.L, 0);
    int maxA = 1;
    for (int l = 0; l < t.L; ++l) {
        const int H = t.H[l], W = t.W[l], A = t.A[l], n0 = t.offs[l];
        if (W > 128 || A > 255) return cudaSuccess;
        maxA = std::max(maxA, A);
        rt_off[l] = (int)(rowtab.size() / 4);
        ct_off[l] = (int)(coltab.size() / 4);
        pl_off[l] = (int)(planes.size() / 4);
        auto at = [&](int r, int c, int a) { return (size_t)(n0 + (r * W + c) * A + a); };
        for (int a = 0; a < A; ++a)
            for (int r = 0; r < H; ++r) {
                const float y0 = cor[at(r, 0, a) * 4 + 0], y1 = cor[at(r, 0, a) * 4 + 2];
                if (!(y1 - y0 > 0.f)) return cudaSuccess;        // the kernel relies on anchors of positive area
                rowtab.insert(rowtab.end(), {y0, y1, y1 - y0, 0.f});
            }
        for (int a = 0; a < A; ++a)
            for (int c = 0; c < W; ++c) {
                const float x0 = cor[at(0, c, a) * 4 + 1], x1 = cor[at(0, c, a) * 4 + 3];
                if (!(x1 - x0 > 0.f)) return cudaSuccess;
                coltab.insert(coltab.end(), {x0, x1, x1 - x0, 0.f});
            }
        for (int a = 0; a < A; ++a) {
            const float* rt = rowtab.data() + ((size_t)rt_off[l] + (size_t)a * H) * 4;
            const float* ct = coltab.data() + ((size_t)ct_off[l] + (size_t)a * W) * 4;
            for (int r = 1; r < H; ++r)
                if (!(rt[r * 4] >= rt[(r - 1) * 4] && rt[r * 4 + 1] >= rt[(r - 1) * 4 + 1])) return cudaSuccess;
            for (int c = 1; c < W; ++c)
                if (!(ct[c * 4] >= ct[(c - 1) * 4] && ct[c * 4 + 1] >= ct[(c - 1) * 4 + 1])) return cudaSuccess;
            std::vector<char> rany(H, 0), cany(W, 0);
            for (int r = 0; r < H; ++r)
                for (int c = 0; c < W; ++c) {
                    const size_t n = at(r, c, a);
                    if (bits(cor[n * 4 + 0]) != bits(rt[r * 4]) || bits(cor[n * 4 + 2]) != bits(rt[r * 4 + 1]) ||
                        bits(cor[n * 4 + 1]) != bits(ct[c * 4]) || bits(cor[n * 4 + 3]) != bits(ct[c * 4 + 1]))
                        return cudaSuccess;
                    if (in[n]) { rany[r] = 1; cany[c] = 1; }
                }
            int r0 = H, r1 = -1, c0 = W, c1 = -1, nr = 0, nc = 0;
            for (int r = 0; r < H; ++r) if (rany[r]) { r0 = std::min(r0, r); r1 = std::max(r1, r); ++nr; }
            for (int c = 0; c < W; ++c) if (cany[c]) { c0 = std::min(c0, c); c1 = std::max(c1, c); ++nc; }
            if (nr && (r1 - r0 + 1 != nr || c1 - c0 + 1 != nc)) return cudaSuccess;
            for (int r = 0; r < H; ++r)
                for (int c = 0; c < W; ++c)
                    if ((in[at(r, c, a)] != 0) != (rany[r] && cany[c])) return cudaSuccess;
            if (!nr) { r0 = 1; r1 = 0; c0 = 1; c1 = 0; }
            planes.insert(planes.end(), {r0, r1, c0, c1});
            max_cells[l] = std::max(max_cells[l], nr * nc);
            n_in[l] += nr * nc;
        }
    }
    // CTA size: the warp count in 6..10 that keeps most warps busy when a warp owns a plane (or one of
    // floor(warps / A) row sub-bands of it), weighted by the inside anchors of the layers swept that way
    int dense_cells = 64;
    if (const char* e = getenv("RONK_ENC_DENSE_CELLS")) dense_cells = atoi(e);       // tuning knob
    auto is_dense = [&](int l) { return max_cells[l] <= dense_cells && t.offs[l + 1] - t.offs[l] <= 8192; };
    int best_nw = 8;
    double best_u = -1.;
    for (int nw = 6; nw <= 10; ++nw) {
        double u = 0., wsum = 0.;
        for (int l = 0; l < t.L; ++l) {
            if (is_dense(l)) continue;
            const int A = t.A[l];
            const double f = A <= nw ? (double)(A * (nw / A)) / nw : (double)A / (((A + nw - 1) / nw) * nw);
            u += f * n_in[l];
            wsum += n_in[l];
        }
        u = wsum > 0 ? u / wsum : 1.;
        if (u >= best_u - 1e-9) { best_u = u; best_nw = nw; }
    }
    if (const char* e = getenv("RONK_ENC_WARPS")) {                                     // tuning knob
        const int v = atoi(e);
        if (v >= 1 && v <= 10) best_nw = v;
    }
    h->grid_threads = 32 * best_nw;

    auto compact_lo = [&](int n) { return (int)(std::lower_bound(idx.begin(), idx.end(), n) - idx.begin()); };
    int targets[3] = {4096, 2048, 1024};                // anchors per row band: coarse / medium / fine cut
    if (const char* e = getenv("RONK_ENC_TARGET")) {    // tuning knob: scales the three cuts
        const int v = atoi(e);
        if (v >= 256 && v <= 8192) { targets[0] = v; targets[1] = v / 2; targets[2] = v / 4; }
    }
    for (int tv = 0; tv < 3; ++tv) {
        std::vector<GridItem> items;
        std::vector<long long> weight;
        for (int l = 0; l < t.L;) {
            GridItem it;
            memset(&it, 0, sizeof(it));
            if (is_dense(l)) {
                // consecutive sparse layers share one dense item
                int l2 = l, nin = 0;
                while (l2 < t.L && is_dense(l2) && t.offs[l2 + 1] - t.offs[l] <= 8192 &&
                       (l2 == l || nin + n_in[l2] <= 1024)) {
                    nin += n_in[l2];
                    ++l2;
                }
                it.mode = 1;
                it.n_lo = t.offs[l];
                it.n_hi = t.offs[l2];
                it.c_lo = compact_lo(it.n_lo);
                it.c_hi = compact_lo(it.n_hi);
                items.push_back(it);
                weight.push_back(it.c_hi - it.c_lo);
                l = l2;
                continue;
            }
            const int H = t.H[l], W = t.W[l], A = t.A[l];
            int rows_max = std::max(1, targets[tv] / (W * A));
            rows_max = std::min(rows_max, 255);
            const int nb = (H + rows_max - 1) / rows_max;
            const int rows = (H + nb - 1) / nb;
            if ((long long)rows * W * A > 8192) return cudaSuccess;
            for (int r_lo = 0; r_lo < H; r_lo += rows) {
                it.mode = 0;
                it.r_lo = r_lo;
                it.rows = std::min(rows, H - r_lo);
                it.n_lo = t.offs[l] + r_lo * W * A;
                it.n_hi = it.n_lo + it.rows * W * A;
                it.H = H; it.W = W; it.A = A;
                it.rt_base = rt_off[l];
                it.ct_base = ct_off[l];
                it.pl_base = pl_off[l];
                it.layer_n0 = t.offs[l];
                it.nsub = std::max(1, std::min(best_nw / A, it.rows));
                int ps = it.rows * W;
                while (ps % 16 != 3) ++ps;               // spreads the planes over the shared-memory banks
                it.pstride = ps;
                it.c_lo = compact_lo(it.n_lo);
                it.c_hi = compact_lo(it.n_hi);
                items.push_back(it);
                weight.push_back(it.c_hi - it.c_lo);
            }
            ++l;
        }
        // heavy items first (the grid is item-major)
        std::vector<int> order(items.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return weight[x] > weight[y]; });
        std::vector<GridItem> sorted;
        size_t smem = 0;
        for (int i : order) {
            const GridItem& g = items[i];
            sorted.push_back(g);
            size_t S = g.mode == 0 ? (size_t)g.A * g.pstride : (size_t)(g.n_hi - g.n_lo);
            S = (S + 1) & ~(size_t)1;
            smem = std::max(smem, S * 8 + (g.mode == 0 ? (size_t)g.A * (g.rows + g.W) * 16 : 0));
        }
        h->n_gitems[tv] = (int)sorted.size();
        h->gitems_smem[tv] = smem;
        cudaError_t e = cudaMalloc(&h->d_gitems[tv], sorted.size() * sizeof(GridItem));
        if (e == cudaSuccess) e = cudaMemcpy(h->d_gitems[tv], sorted.data(), sorted.size() * sizeof(GridItem), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaMalloc(&h->d_rowtab, rowtab.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_coltab, coltab.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_planes, planes.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_rowtab, rowtab.data(), rowtab.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_coltab, coltab.data(), coltab.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_planes, planes.data(), planes.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    h->grid_ok = 1;
    return cudaSuccess;
}

cudaError_t finish_compaction(ronk_anchors* h) {
    const int N = h->tab.N;
    std::vector<uint8_t> in(N);
    std::vector<float> cor((size_t)N * 4);
    cudaError_t e = cudaMemcpy(in.data(), h->d_inside, (size_t)N, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(cor.data(), h->d_cor, (size_t)N * 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    std::vector<int> idx, cidx(N);
    std::vector<float> ccor;
    idx.reserve(N);
    for (int n = 0; n < N; ++n) {
        if (in[n]) {
            cidx[n] = (int)idx.size();
            idx.push_back(n);
            ccor.insert(ccor.end(), cor.begin() + (size_t)n * 4, cor.begin() + (size_t)n * 4 + 4);
        } else {
            cidx[n] = -1;
        }
    }
    h->n_inside = (int)idx.size();
    h->anchors_nice = 1;
    for (float v : ccor) {
        float m = fabsf(v);
        if (!(v == 0.f || (m >= 3.0517578125e-05f && m <= 32768.f))) h->anchors_nice = 0;
    }
    const size_t nin = idx.size() ? idx.size() : 1;
    e = cudaMalloc(&h->d_inside_idx, nin * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_cidx, (size_t)N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_ccor, nin * 16);
    if (e == cudaSuccess && !idx.empty()) e = cudaMemcpy(h->d_inside_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_cidx, cidx.data(), (size_t)N * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !idx.empty()) e = cudaMemcpy(h->d_ccor, ccor.data(), ccor.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;

    e = build_grid(h, cor, in, idx);
    if (e != cudaSuccess) return e;

    // ---- post-process tile table (see common.cuh)
    {
        const LayerTable& t = h->tab;
        std::vector<int> plain;
        for (int l = 0; l < t.L; ++l) {
            const int n_l = t.offs[l + 1] - t.offs[l];
            for (int t0 = 0; t0 < n_l; t0 += kTileRows) {
                const int rows = n_l - t0 < kTileRows ? n_l - t0 : kTileRows;
                plain.insert(plain.end(), {l, t0, rows, n_l});
            }
        }
        const int tpi = (int)(plain.size() / 4);
        int m = (int)(0.618 * tpi);
        auto gcd = [](int a, int b) { while (b) { int q = a % b; a = b; b = q; } return a; };
        if (m < 1) m = 1;
        while (m < tpi && gcd(m, tpi) != 1) ++m;
        if (m >= tpi) m = 1;
        std::vector<int> perm((size_t)tpi * 4);
        for (int rp = 0; rp < tpi; ++rp) {
            const int r = (int)(((long long)rp * m) % tpi);
            for (int q = 0; q < 4; ++q) perm[(size_t)rp * 4 + q] = plain[(size_t)r * 4 + q];
        }
        h->tiles_per_image = tpi;
        h->tile_perm_mul = m;
        e = cudaMalloc(&h->d_tile_tab, perm.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_tile_tab, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
    }

    // ---- work-item tables (see common.cuh).  A set of 64 consecutive inside anchors is "heavy" when its
    // bounding extent, clipped to the image, covers more than 40 % of it: then nearly every GT box
    // survives the cull and the item is as long as the GT list.
    const int Nin = h->n_inside;
    const int nsets = (Nin + 63) / 64;
    std::vector<char> heavy(nsets > 0 ? nsets : 1, 0);
    for (int sidx = 0; sidx < nsets; ++sidx) {
        float y0 = 1e30f, x0 = 1e30f, y1 = -1e30f, x1 = -1e30f;
        for (int c = sidx * 64; c < Nin && c < sidx * 64 + 64; ++c) {
            y0 = fminf(y0, ccor[(size_t)c * 4 + 0]); x0 = fminf(x0, ccor[(size_t)c * 4 + 1]);
            y1 = fmaxf(y1, ccor[(size_t)c * 4 + 2]); x1 = fmaxf(x1, ccor[(size_t)c * 4 + 3]);
        }
        float hh = fminf(y1, 1.f) - fmaxf(y0, 0.f), ww = fminf(x1, 1.f) - fmaxf(x0, 0.f);
        heavy[sidx] = (hh > 0.f && ww > 0.f && hh * ww > 0.4f) ? 1 : 0;
    }
    auto flat_lo = [&](int c0) { return c0 == 0 ? 0 : idx[c0]; };
    auto flat_hi = [&](int c1) { return c1 >= Nin ? N : idx[c1]; };
    for (int t = 0; t < 3; ++t) {
        std::vector<int> items;
        for (int s0 = 0; s0 < (nsets > 0 ? nsets : 1); s0 += 4) {
            int nh = 0;
            for (int q = s0; q < s0 + 4 && q < nsets; ++q) nh += heavy[q];
            const bool cut = (t == 2) || (t == 1 && nh >= 2);
            if (cut) {
                for (int q = s0; q < s0 + 4 && q < (nsets > 0 ? nsets : 1); ++q) {
                    int c0 = q * 64, c1 = c0 + 64;
                    items.insert(items.end(), {c0, flat_lo(c0), flat_hi(c1), 4});
                }
            } else {
                int c0 = s0 * 64, c1 = c0 + 256;
                items.insert(items.end(), {c0, flat_lo(c0), flat_hi(c1), 1});
            }
        }
        h->n_items[t] = (int)(items.size() / 4);
        e = cudaMalloc(&h->d_items[t], items.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_items[t], items.data(), items.size() * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
    }
    return e;
}

}  // namespace ronk

using namespace ronk;

extern "C" int ronk_anchors_create_flat(int img_h, int img_w, int N, const float* yxhw, const int* border,
                                        ronk_anchors_t** out) {
    RONK_REQUIRE(out != nullptr, RONK_EINVAL, "ronk_anchors_create_flat: out is NULL");
    *out = nullptr;
    RONK_REQUIRE(yxhw && N >= 1 && N <= (1 << 24) && img_h > 0 && img_w > 0, RONK_EINVAL,
                 "ronk_anchors_create_flat: bad argument");
    ronk_anchors* h = new (std::nothrow) ronk_anchors();
    RONK_REQUIRE(h != nullptr, RONK_ENOMEM, "ronk_anchors_create_flat: out of host memory");
    memset(h, 0, sizeof(*h));
    h->kind = -1;
    h->img_h = img_h;
    h->img_w = img_w;
    h->tab.L = 1;
    h->tab.N = N;
    h->tab.offs[0] = 0;
    h->tab.offs[1] = N;
    h->tab.H[0] = N;
    h->tab.W[0] = 1;
    h->tab.A[0] = 1;
    int* d_border = nullptr;
    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_dec, (size_t)N * 16);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_enc, (size_t)N * 16);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_cor, (size_t)N * 16);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_inside, (size_t)N);
    if (e == cudaSuccess && border) e = cudaMalloc(&d_border, (size_t)N * 4);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_dec, yxhw, (size_t)N * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_enc, yxhw, (size_t)N * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && border) e = cudaMemcpy(d_border, border, (size_t)N * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        anchor_flat_kernel<<<(N + 255) / 256, 256>>>((const float4*)h->d_enc, d_border, N, img_h, img_w,
                                                     (float4*)h->d_cor, h->d_inside);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = finish_compaction(h);
    if (d_border) cudaFree(d_border);
    if (e != cudaSuccess) {
        int r = cuda_fail(e, "ronk_anchors_create_flat");
        ronk_anchors_destroy(h);
        return (e == cudaErrorMemoryAllocation) ? RONK_ENOMEM : r;
    }
    *out = h;
    return RONK_OK;
}

extern "C" int ronk_anchors_create(int kind, int img_h, int img_w, int num_layers,
                                   const int* feat_shapes, const double* sizes, const int* n_sizes,
                                   const double* ratios, const int* n_ratios, const double* steps,
                                   double offset, const int* borders, ronk_anchors_t** out) {
    RONK_REQUIRE(out != nullptr, RONK_EINVAL, "ronk_anchors_create: out is NULL");
    *out = nullptr;
    RONK_REQUIRE(kind == RONK_KIND_RON || kind == RONK_KIND_SSD, RONK_EINVAL, "ronk_anchors_create: unknown kind");
    RONK_REQUIRE(num_layers >= 1 && num_layers <= kMaxLayers, RONK_ELIMIT, "ronk_anchors_create: 1..16 layers supported");
    RONK_REQUIRE(img_h > 0 && img_w > 0, RONK_EINVAL, "ronk_anchors_create: bad image shape");
    RONK_REQUIRE(feat_shapes && sizes && n_sizes && ratios && n_ratios && steps, RONK_EINVAL,
                 "ronk_anchors_create: NULL parameter array");

    ronk_anchors* h = new (std::nothrow) ronk_anchors();
    RONK_REQUIRE(h != nullptr, RONK_ENOMEM, "ronk_anchors_create: out of host memory");
    memset(h, 0, sizeof(*h));
    AnchorGenParams* p = new (std::nothrow) AnchorGenParams();
    if (!p) { delete h; set_error("ronk_anchors_create: out of host memory"); return RONK_ENOMEM; }
    memset(p, 0, sizeof(*p));
    h->kind = kind;
    h->img_h = img_h;
    h->img_w = img_w;
    LayerTable& t = h->tab;
    t.L = num_layers;
    int n = 0, shapes = 0, so = 0, ro = 0;
    int rc = RONK_OK;
    for (int l = 0; l < num_layers && rc == RONK_OK; ++l) {
        int S = n_sizes[l], R = n_ratios[l];
        int A = (kind == RONK_KIND_RON) ? S * R : S + R;
        if (S < 1 || R < 0 || A < 1 || shapes + A > kMaxShapes || (kind == RONK_KIND_SSD && S > 2)) {
            set_error("ronk_anchors_create: unsupported sizes/ratios for a layer");
            rc = RONK_ELIMIT;
            break;
        }
        t.offs[l] = n;
        t.H[l] = feat_shapes[2 * l];
        t.W[l] = feat_shapes[2 * l + 1];
        t.A[l] = A;
        t.hw_off[l] = shapes;
        if (t.H[l] < 1 || t.W[l] < 1) { set_error("ronk_anchors_create: bad feature shape"); rc = RONK_EINVAL; break; }
        const double* sz = sizes + so;
        const double* rt = ratios + ro;
        float* hh = h->h_host + shapes;
        float* ww = h->w_host + shapes;
        if (kind == RONK_KIND_RON) {
            // a = ratio_index * S + size_index; h = s / img_h / sqrt(r), w = s / img_w * sqrt(r) in double
            for (int i = 0; i < R; ++i)
                for (int j = 0; j < S; ++j) {
                    hh[i * S + j] = (float)(sz[j] / (double)img_h / sqrt(rt[i]));
                    ww[i * S + j] = (float)(sz[j] / (double)img_w * sqrt(rt[i]));
                }
        } else {
            hh[0] = (float)(sz[0] / (double)img_h);
            ww[0] = (float)(sz[0] / (double)img_w);
            int di = 1;
            if (S > 1) {
                hh[1] = (float)(sqrt(sz[0] * sz[1]) / (double)img_h);
                ww[1] = (float)(sqrt(sz[0] * sz[1]) / (double)img_w);
                di = 2;
            }
            for (int i = 0; i < R; ++i) {
                hh[i + di] = (float)(sz[0] / (double)img_h / sqrt(rt[i]));
                ww[i + di] = (float)(sz[0] / (double)img_w * sqrt(rt[i]));
            }
        }
        p->step[l] = (float)steps[l];
        if (borders) {
            double b = (double)borders[l];
            p->lo_y[l] = (float)(-b * 1. / img_h);
            p->lo_x[l] = (float)(-b * 1. / img_w);
            p->hi_y[l] = (float)((img_h + b) * 1. / img_h);
            p->hi_x[l] = (float)((img_w + b) * 1. / img_w);
        }
        long long nl = (long long)t.H[l] * t.W[l] * A;
        if (n + nl > (1 << 24)) { set_error("ronk_anchors_create: more than 2^24 anchors"); rc = RONK_ELIMIT; break; }
        n += (int)nl;
        shapes += A;
        so += S;
        ro += R;
    }
    if (rc != RONK_OK) { delete p; delete h; return rc; }
    t.offs[num_layers] = n;
    t.N = n;
    p->tab = t;
    memcpy(p->h, h->h_host, sizeof(p->h));
    memcpy(p->w, h->w_host, sizeof(p->w));
    p->img_h = (float)img_h;
    p->img_w = (float)img_w;
    p->offset = (float)offset;
    p->has_border = borders ? 1 : 0;

    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_dec, (size_t)n * 16);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_enc, (size_t)n * 16);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_cor, (size_t)n * 16);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_inside, (size_t)n);
    if (e == cudaSuccess) {
        anchor_gen_kernel<<<(n + 255) / 256, 256>>>(*p, (float4*)h->d_dec, (float4*)h->d_enc,
                                                    (float4*)h->d_cor, h->d_inside);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();   // creation only; compute calls never sync
    if (e == cudaSuccess) e = finish_compaction(h);
    delete p;
    if (e != cudaSuccess) {
        int r = cuda_fail(e, "ronk_anchors_create");
        ronk_anchors_destroy(h);
        return (e == cudaErrorMemoryAllocation) ? RONK_ENOMEM : r;
    }
    *out = h;
    return RONK_OK;
}

extern "C" void ronk_anchors_destroy(ronk_anchors_t* h) {
    if (!h) return;
    if (h->d_dec) cudaFree(h->d_dec);
    if (h->d_enc) cudaFree(h->d_enc);
    if (h->d_cor) cudaFree(h->d_cor);
    if (h->d_inside) cudaFree(h->d_inside);
    if (h->d_inside_idx) cudaFree(h->d_inside_idx);
    if (h->d_cidx) cudaFree(h->d_cidx);
    if (h->d_ccor) cudaFree(h->d_ccor);
    for (int t = 0; t < 3; ++t)
        if (h->d_items[t]) cudaFree(h->d_items[t]);
    if (h->d_tile_tab) cudaFree(h->d_tile_tab);
    if (h->d_rowtab) cudaFree(h->d_rowtab);
    if (h->d_coltab) cudaFree(h->d_coltab);
    if (h->d_planes) cudaFree(h->d_planes);
    for (int t = 0; t < 3; ++t)
        if (h->d_gitems[t]) cudaFree(h->d_gitems[t]);
    delete h;
}

extern "C" int ronk_anchors_num(const ronk_anchors_t* h) { return h ? h->tab.N : RONK_EINVAL; }
extern "C" int ronk_anchors_num_layers(const ronk_anchors_t* h) { return h ? h->tab.L : RONK_EINVAL; }

extern "C" int ronk_anchors_layer_info(const ronk_anchors_t* h, int layer, int* H, int* W, int* A, int* offset) {
    RONK_REQUIRE(h && layer >= 0 && layer < h->tab.L, RONK_EINVAL, "ronk_anchors_layer_info: bad handle or layer");
    if (H) *H = h->tab.H[layer];
    if (W) *W = h->tab.W[layer];
    if (A) *A = h->tab.A[layer];
    if (offset) *offset = h->tab.offs[layer];
    return RONK_OK;
}

extern "C" const void* ronk_anchors_table(const ronk_anchors_t* h, int which) {
    if (!h) return nullptr;
    switch (which) {
        case 0: return h->d_dec;
        case 1: return h->d_enc;
        case 2: return h->d_cor;
        case 3: return h->d_inside;
    }
    return nullptr;
}

extern "C" int ronk_anchors_layer_hw(const ronk_anchors_t* h, int layer, float* hh, float* ww) {
    RONK_REQUIRE(h && layer >= 0 && layer < h->tab.L && hh && ww, RONK_EINVAL, "ronk_anchors_layer_hw: bad argument");
    int A = h->tab.A[layer], o = h->tab.hw_off[layer];
    memcpy(hh, h->h_host + o, sizeof(float) * A);
    memcpy(ww, h->w_host + o, sizeof(float) * A);
    return RONK_OK;
}
