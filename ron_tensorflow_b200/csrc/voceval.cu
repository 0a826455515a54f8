// PASCAL VOC detection matching of the reference's offline evaluator:
// datasets/voc_eval.py:249-281.  One class at a time; the detections arrive sorted by decreasing confidence
// and grouped by image (the "already detected" flags only couple detections of one image, so images are
// independent): one warp per image walks its detections in order, the lanes take the image's ground-truth
// boxes of the class.  Everything is float64 like the reference (pixel coordinates parsed from the result
// files), one rounding per operation:
//   inters = max(ixmax - ixmin, 0) * max(iymax - iymin, 0);  uni = (area_det + area_gt) - inters
//   ovmax = max(inters / uni), jmax = first arg-max; a NaN overlap makes the maximum NaN (no match)
//   ovmax > ovthresh: difficult -> neither; first hit -> TP, later hits -> FP;  otherwise FP.
#include <math_constants.h>

#include "common.cuh"

namespace ronk {

__global__ void __launch_bounds__(32)
voc_match_kernel(const double4* __restrict__ det, const int* __restrict__ det_off, const double4* __restrict__ gt,
                 const int* __restrict__ gt_off, const uint8_t* __restrict__ gt_difficult, double ovthresh,
                 uint8_t* __restrict__ tp, uint8_t* __restrict__ fp) {
    extern __shared__ unsigned char s_det[];                       // "det" flag of every GT box of the image
    const unsigned full = 0xffffffffu;
    const int img = blockIdx.x, lane = threadIdx.x;
    const int d0 = det_off[img], d1 = det_off[img + 1], g0 = gt_off[img], g1 = gt_off[img + 1];
    const int ng = g1 - g0;
    for (int j = lane; j < ng; j += 32) s_det[j] = 0;
    __syncwarp();
    for (int d = d0; d < d1; ++d) {
        const double4 bb = det[d];                                 // x1 y1 x2 y2
        const double area = (bb.z - bb.x) * (bb.w - bb.y);
        double best = -CUDART_INF;                                 // ovmax = -inf without ground truth (:250)
        int arg = 0x7fffffff;
        bool nan = false;
        for (int j = lane; j < ng; j += 32) {
            const double4 g = gt[g0 + j];
            const double iw = fmax(fmin(g.z, bb.z) - fmax(g.x, bb.x), 0.);
            const double ih = fmax(fmin(g.w, bb.w) - fmax(g.y, bb.y), 0.);
            const double inters = iw * ih;
            const double uni = (area + (g.z - g.x) * (g.w - g.y)) - inters;
            const double ov = inters / uni;
            nan = nan || ov != ov;
            if (ov > best) { best = ov; arg = j; }                 // ascending j: first arg-max per lane
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(full, best, o);
            const int oa = __shfl_xor_sync(full, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        nan = __any_sync(full, nan);
        if (lane == 0) {
            uint8_t t = 0, f = 0;
            if (!nan && best > ovthresh) {                         // :270
                if (!gt_difficult[g0 + arg]) {
                    if (!s_det[arg]) { t = 1; s_det[arg] = 1; }
                    else f = 1;
                }
            } else {
                f = 1;
            }
            tp[d] = t;
            fp[d] = f;
        }
        __syncwarp();
    }
}

}  // namespace ronk

using namespace ronk;

extern "C" int ronk_voc_match(const double* det_boxes, const int32_t* det_offsets, const double* gt_boxes,
                              const int32_t* gt_offsets, const uint8_t* gt_difficult, int n_images, int max_gt,
                              double ovthresh, uint8_t* out_tp, uint8_t* out_fp, void* stream) {
    RONK_REQUIRE(n_images >= 0 && max_gt >= 0 && det_offsets && gt_offsets, RONK_EINVAL, "ronk_voc_match: bad argument");
    if (n_images == 0) return RONK_OK;
    RONK_REQUIRE(((uintptr_t)det_boxes % 32) == 0 && ((uintptr_t)gt_boxes % 32) == 0, RONK_EINVAL,
                 "ronk_voc_match: box pointers must be 32-byte aligned");
    RONK_REQUIRE(max_gt <= 48 * 1024, RONK_ELIMIT, "ronk_voc_match: more than 49152 ground-truth boxes in one image");
    voc_match_kernel<<<(unsigned)n_images, 32, (size_t)(max_gt > 0 ? max_gt : 1), (cudaStream_t)stream>>>(
        (const double4*)det_boxes, det_offsets, (const double4*)gt_boxes, gt_offsets, gt_difficult, ovthresh, out_tp,
        out_fp);
    RONK_LAUNCHED();
    return RONK_OK;
}
