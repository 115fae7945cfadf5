/*
 * ctd_b200.h -- C ABI of libctd_b200.so, the B200 (sm_100a) implementation of the
 * per-pixel custom ops of "Connecting the Dots" (reference: torchext/ext/).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Each entry
 * point names the reference interface it replaces (paths relative to the reference
 * repository).  Two families:
 *
 *   ctd_<op>_<dtype>(..., stream)      DEVICE pointers; asynchronous launch on `stream`
 *                                      (cudaStream_t; NULL = legacy default stream);
 *                                      what ext_cuda.cpp's *_cuda / photometric_loss_*
 *                                      functions would call instead of ext_kernel.cu.
 *   ctd_host_<op>_<dtype>(...)         HOST pointers (pageable or pinned); copies in,
 *                                      runs the same kernels, copies out, returns when the
 *                                      result is in host memory; what ext_cpu.cpp's *_cpu
 *                                      functions would call instead of iterate_cpu.
 *
 * Every function returns CTD_OK (0) or a CTD_ERR_* code and never terminates the process
 * (the reference's CUDA_CHECK calls exit(-1), common_cuda.h:11-20).  ctd_last_error()
 * returns a thread-local message for the last failure.  There is no CPU fallback: without
 * a CUDA device every compute entry point returns CTD_ERR_CUDA.
 *
 * Layouts are the reference's: contiguous row-major, es/ta/grad [B,C,H,W], out [B,1,H,W],
 * xyz [B,H,W,3], K [3,3], cost volume [D,H,W] per image.  Index outputs are int64,
 * CrossCheck masks are uint8 (ext_cpu.cpp:31,50,76).
 */
#ifndef CTD_B200_H
#define CTD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ctd_stream_t; /* a cudaStream_t */

enum {
  CTD_OK = 0,
  CTD_ERR_INVALID = 1, /* bad argument (shape, type id, null pointer) */
  CTD_ERR_CUDA = 2,    /* CUDA runtime error, message in ctd_last_error() */
  CTD_ERR_NOMEM = 3    /* workspace allocation failed */
};

/* loss type ids, torchext/ext/ext.h:196-199 */
enum { CTD_LOSS_MSE = 0, CTD_LOSS_SAD = 1, CTD_LOSS_CENSUS_MSE = 2, CTD_LOSS_CENSUS_SAD = 3 };

const char* ctd_last_error(void);
/* library version and the SM architecture the kernels were compiled for ("sm_100a") */
const char* ctd_version(void);
/* number of kernel launches issued by this library in the calling process so far */
uint64_t ctd_launch_count(void);
/* test/tuning switches: "force_generic" = 1 routes every op through its generic kernel */
int ctd_set_option(const char* name, int value);

/* ---- PhotometricLossForward: ext.h:201-266, ext_cuda.cpp:92-104 (photometric_loss_forward) */
int ctd_photometric_fwd_f32(const float* es, const float* ta, float* out, int64_t B, int64_t C,
                            int64_t H, int64_t W, int block_size, int type, float eps,
                            ctd_stream_t stream);
int ctd_photometric_fwd_f64(const double* es, const double* ta, double* out, int64_t B, int64_t C,
                            int64_t H, int64_t W, int block_size, int type, float eps,
                            ctd_stream_t stream);
/* ---- PhotometricLossBackward: ext.h:268-344, ext_cuda.cpp:109-123 (photometric_loss_backward).
 * grad_in [B,C,H,W] is fully overwritten (no zero-initialisation needed, no atomics). */
int ctd_photometric_bwd_f32(const float* es, const float* ta, const float* grad_out, float* grad_in,
                            int64_t B, int64_t C, int64_t H, int64_t W, int block_size, int type,
                            float eps, ctd_stream_t stream);
int ctd_photometric_bwd_f64(const double* es, const double* ta, const double* grad_out,
                            double* grad_in, int64_t B, int64_t C, int64_t H, int64_t W,
                            int block_size, int type, float eps, ctd_stream_t stream);

/* ---- forward + backward in one call (ext.h:201-344), for callers whose grad_out does not depend on the loss
 * map -- the reference's only caller builds it from the mask alone (model/networks.py:377).  The census modes
 * run a single fused kernel (the backward's window terms already contain the forward's); results are identical
 * to calling ctd_photometric_fwd_f32 then ctd_photometric_bwd_f32. */
int ctd_photometric_fwd_bwd_f32(const float* es, const float* ta, const float* grad_out, float* out,
                                float* grad_in, int64_t B, int64_t C, int64_t H, int64_t W, int block_size,
                                int type, float eps, ctd_stream_t stream);

/* The same plus the caller's masked-mean terms (model/networks.py:377) from the same pass:
 * sums2[0] = sum(mask * out), sums2[1] = sum(mask), deterministic.  mask [B,1,H,W]. */
int ctd_photometric_fwd_bwd_masked_f32(const float* es, const float* ta, const float* grad_out, const float* mask,
                                       float* out, float* grad_in, float* sums2, int64_t B, int64_t C, int64_t H,
                                       int64_t W, int block_size, int type, float eps, ctd_stream_t stream);

/* ---- the disparity warp in front of the loss: model/networks.py:362-371 (RectifiedPatternSimilarityLoss.tforward),
 * grid_sample(pattern, grid(disp), bilinear, padding_mode='border', align_corners=False) with the reference's (W-1)
 * grid normalisation.  pattern [Bp,1,Hp,Wp] (Bp = 1 or B), disp [B,1,H,W] -> out [B,1,H,W]; backward: gradient
 * w.r.t. disp (the pattern is a constant of the model). */
int ctd_warp_pattern_fwd_f32(const float* pattern, const float* disp, float* out, int64_t B, int64_t Bp, int64_t Hp,
                             int64_t Wp, int64_t H, int64_t W, ctd_stream_t stream);
int ctd_warp_pattern_bwd_f32(const float* pattern, const float* disp, const float* grad_out, float* grad_disp,
                             int64_t B, int64_t Bp, int64_t Hp, int64_t Wp, int64_t H, int64_t W, ctd_stream_t stream);

/* ---- geometric loss of the stage-2 trainer, one direction: model/networks.py:474-498
 * (ProjectionDepthSimilarityLoss.fwd over ProjectionBaseLoss.unproject/transform/project, networks.py:436-472).
 * depthA, depthB [B,1,H,W]; ray [H*W,3] (the module's uv @ Ki^T table); K [3,3]; RA, RB [B,3,3]; tA, tB [B,3];
 * all DEVICE pointers.  sums2[0] = sum |d - grid_sample(depthB)| (clamped to [0, clamp] when clamp > 0),
 * sums2[1] = B*H*W, deterministic.  Gradients of sums2[0] * scale: grad_depthA [B,1,H,W] is written
 * (direct_accumulate = 0) or added to (1); grad_depthB receives the bilinear scatter by atomicAdd and must be
 * initialised by the caller.  Either gradient pointer may be NULL. */
int ctd_depth_similarity_f32(const float* depthA, const float* depthB, const float* ray, const float* K,
                             const float* RA, const float* tA, const float* RB, const float* tB,
                             float* grad_depthA, float* grad_depthB, float* sums2, int64_t B, int64_t H, int64_t W,
                             float clamp, float scale, int direct_accumulate, ctd_stream_t stream);

/* ---- disparity (smoothness / edge) loss of the stage-1 trainer: model/networks.py:380-412 (DisparityLoss.tforward)
 * over the 5x5 Sobel filter with replicate padding (networks.py:537-565).  disp, edge, grad_disp, grad_edge
 * [B,1,H,W], DEVICE pointers; edge may be NULL (the reference's edge=None branch: mean of the clamped gradient
 * magnitude).  sums2[0] = sum of the per-pixel loss, sums2[1] = B*H*W, deterministic; the gradients are those of
 * sums2[0] * scale, fully overwritten, no atomics.  grad_disp / grad_edge may be NULL. */
int ctd_disparity_loss_f32(const float* disp, const float* edge, float* grad_disp, float* grad_edge, float* sums2,
                           int64_t B, int64_t H, int64_t W, float scale, ctd_stream_t stream);

/* ---- XCorrVolFunctor: ext.h:120-191, ext_cuda.cpp:73-86 (xcorrvol_cuda).  The reference has no
 * batch dimension; here in0,in1 are [B,C,H,W] and out is [B,D,H,W] (B=1 is the reference call). */
int ctd_xcorrvol_f32(const float* in0, const float* in1, float* out, int64_t B, int64_t C, int64_t H,
                     int64_t W, int64_t n_disps, int block_size, ctd_stream_t stream);
int ctd_xcorrvol_f64(const double* in0, const double* in1, double* out, int64_t B, int64_t C,
                     int64_t H, int64_t W, int64_t n_disps, int block_size, ctd_stream_t stream);

/* ---- ProjNNFunctor: ext.h:65-117, ext_cuda.cpp:47-69 (proj_nn_cuda).  K is a DEVICE pointer to
 * 9 values (row-major 3x3), like the reference's K tensor. */
int ctd_proj_nn_f32(const float* xyz0, const float* xyz1, const float* K, int64_t* out, int64_t B,
                    int64_t H, int64_t W, int patch_size, ctd_stream_t stream);
int ctd_proj_nn_f64(const double* xyz0, const double* xyz1, const double* K, int64_t* out, int64_t B,
                    int64_t H, int64_t W, int patch_size, ctd_stream_t stream);

/* ---- NNFunctor<T,3>: ext.h:13-46, ext_cuda.cpp:9-26 (nn_cuda).  in0 [N0,3], in1 [N1,3]. */
int ctd_nn_f32(const float* in0, const float* in1, int64_t* out, int64_t N0, int64_t N1,
               ctd_stream_t stream);
int ctd_nn_f64(const double* in0, const double* in1, int64_t* out, int64_t N0, int64_t N1,
               ctd_stream_t stream);

/* ---- CrossCheckFunctor: ext.h:48-63, ext_cuda.cpp:31-43 (crosscheck_cuda).  N1 is used only to
 * reject out-of-range gathers (they yield 0; the reference reads out of bounds there). */
int ctd_crosscheck(const int64_t* in0, const int64_t* in1, uint8_t* out, int64_t N0, int64_t N1,
                   ctd_stream_t stream);

/* ---- LCN: model/networks.py:507-533 (LCN.tforward).  x [N,1,H,W] -> lcn, std [N,1,H,W]. */
int ctd_lcn_f32(const float* x, float* lcn, float* std, int64_t N, int64_t H, int64_t W, int radius,
                float epsilon, ctd_stream_t stream);
int ctd_lcn_f64(const double* x, double* lcn, double* std, int64_t N, int64_t H, int64_t W,
                int radius, double epsilon, ctd_stream_t stream);

/* ---- RectifiedPatternSimilarityLoss.tforward + backward to the disparity, model/networks.py:358-378, as ONE kernel
 * (census modes, block 9): grid_sample(pattern, grid(disp), border) formed inside the loss kernel's tile loader, loss map,
 * masked-mean terms sums2 = (sum(mask * out), sum(mask)) and d loss / d disp for `grad_out` (w.r.t. the loss map).
 * pattern [Bp,1,Hp,Wp] with Bp = 1 or B; everything else [B,1,H,W]. */
int ctd_pattern_similarity_f32(const float* pattern, const float* disp, const float* ta, const float* grad_out,
                               const float* mask, float* pattern_proj, float* out, float* grad_disp, float* sums2,
                               int64_t B, int64_t Bp, int64_t Hp, int64_t Wp, int64_t H, int64_t W, int type, float eps,
                               ctd_stream_t stream);

/* ---- LCN backward: what autograd produces for networks.LCN (model/networks.py:523-533) w.r.t. its input, given the
 * forward's input x, its outputs (lcn, std) and upstream gradients for both outputs (either may be NULL = zero). */
int ctd_lcn_bwd_f32(const float* x, const float* lcn, const float* std, const float* grad_lcn, const float* grad_std,
                    float* grad_x, int64_t N, int64_t H, int64_t W, int radius, float epsilon, ctd_stream_t stream);

/* ---- the data generator's offline LCN: data/lcn/lcn.pyx:16-58 `normalize(img, kernel_size, epsilon)` (called at
 * data/create_syn_data.py:182) for a batch of B images [M,N] -- no padding (a border of kernel_size pixels stays 0),
 * centred two-pass variance, std = sqrt(var); bit-identical to the Cython build. */
int ctd_lcn_cython_f32(const float* img, float* lcn, float* std, int64_t B, int64_t M, int64_t N, int kernel_size,
                       float epsilon, ctd_stream_t stream);

/* ---- masked loss reduction of the caller, model/networks.py:377: val = (mask*diff).sum() / mask.sum().
 * out2[0] = sum(mask*diff), out2[1] = sum(mask), deterministic.  `workspace`: device memory of
 * ctd_masked_sums_workspace_bytes() bytes, zero-filled once before first use, one per concurrent stream. */
int64_t ctd_masked_sums_workspace_bytes(void);
int ctd_masked_sums_f32(const float* diff, const float* mask, int64_t n, float* out2, void* workspace,
                        ctd_stream_t stream);

/* ---- host-buffer entry points (fp32): same arguments, HOST pointers, synchronous.  They stage
 * through a per-thread grow-only device workspace on the current device. */
int ctd_host_photometric_fwd_f32(const float* es, const float* ta, float* out, int64_t B, int64_t C,
                                 int64_t H, int64_t W, int block_size, int type, float eps);
int ctd_host_photometric_bwd_f32(const float* es, const float* ta, const float* grad_out,
                                 float* grad_in, int64_t B, int64_t C, int64_t H, int64_t W,
                                 int block_size, int type, float eps);
/* forward and backward in one staged call: es/ta cross the bus once */
int ctd_host_photometric_fwd_bwd_f32(const float* es, const float* ta, const float* grad_out,
                                     float* out, float* grad_in, int64_t B, int64_t C, int64_t H,
                                     int64_t W, int block_size, int type, float eps);
/* the whole use the reference's caller makes of the loss (model/networks.py:376-377) in one staged call: loss map
 * (`out`, may be NULL: then it is not downloaded), d loss / d es for the caller's grad_out, and the masked-mean terms
 * sums2[0] = sum(mask * loss), sums2[1] = sum(mask) (host float[2]) */
int ctd_host_photometric_fwd_bwd_masked_f32(const float* es, const float* ta, const float* grad_out,
                                            const float* mask, float* out, float* grad_in, float* sums2,
                                            int64_t B, int64_t C, int64_t H, int64_t W, int block_size,
                                            int type, float eps);
int ctd_host_xcorrvol_f32(const float* in0, const float* in1, float* out, int64_t B, int64_t C,
                          int64_t H, int64_t W, int64_t n_disps, int block_size);
int ctd_host_proj_nn_f32(const float* xyz0, const float* xyz1, const float* K, int64_t* out,
                         int64_t B, int64_t H, int64_t W, int patch_size);
int ctd_host_nn_f32(const float* in0, const float* in1, int64_t* out, int64_t N0, int64_t N1);
int ctd_host_crosscheck(const int64_t* in0, const int64_t* in1, uint8_t* out, int64_t N0, int64_t N1);
int ctd_host_lcn_f32(const float* x, float* lcn, float* std, int64_t N, int64_t H, int64_t W,
                     int radius, float epsilon);
/* Deferred mode for the calling thread: between begin and end the ctd_host_* calls only enqueue their copies and
 * kernels (they return at once), so consecutive calls overlap on the bus; every result is in host memory when
 * ctd_host_end_batch() returns.  Input buffers must stay untouched, output buffers unread, until then.
 * An input that several photometric calls of one batch read (same host address and size: the image pair of two
 * loss types, the gradient weights) is uploaded once.  An input that is an OUTPUT of an earlier call of the same batch
 * (LCN's std used as the loss mask, ProjNN's indices fed to CrossCheck) is taken from that call's device buffer when
 * address and size match exactly (no copy at all); any other overlap with an output still in flight first waits for
 * the downloads, then uploads what the host holds. */
int ctd_host_begin_batch(void);
int ctd_host_end_batch(void);
/* Two batches in flight: ctd_host_end_batch_async() returns once the batch is enqueued, ctd_host_wait_batch() waits for the
 * OLDEST batch not waited for yet (CTD_OK at once if there is none); a thread may have two such batches outstanding -- the
 * third ctd_host_begin_batch waits for the oldest itself.  With steps issued as begin / calls / end_async / wait (for
 * the previous step), the uploads of step k + 1 cross the bus under the downloads of step k.  Results of a batch are in
 * host memory after its wait; until then its host inputs must not change, its host outputs must not be read -- or be
 * passed as inputs to a later batch.  ctd_host_end_batch() (and any ctd_host_* call outside a batch) waits for
 * everything outstanding. */
int ctd_host_end_batch_async(void);
int ctd_host_wait_batch(void);
/* host-to-device bytes the calling thread's current (or last) batch copied, and bytes it did not have to copy again */
void ctd_host_batch_stats(uint64_t* h2d_bytes, uint64_t* h2d_bytes_saved);
/* Repeated batches (opt-in: ctd_set_option("host_graphs", 1)): a batch whose calls (entry points, arguments, host
 * addresses) equal those of an earlier batch of the thread, with every host buffer in pinned memory, is captured into
 * a CUDA graph the second time and replayed from the third on -- ctd_host_end_batch then issues ONE graph launch
 * instead of the batch's copies, kernels and events.  The calls of a replayed batch return at once; host inputs are
 * read when ctd_host_end_batch runs (the contract above: they must not change while the batch is open).  A batch that
 * stops matching is issued the ordinary way from the point of divergence.  Measured on the bench's step: the same
 * step time (the step is bound by the bus), the host thread's time in the calls drops from 0.09 ms to 0.01 ms (profiles/r02_e2e_chunks.json).
 * Counters for the calling thread: graphs captured, graph launches (the capturing batch included), expectations given up. */
void ctd_host_graph_stats(uint64_t* captured, uint64_t* launched, uint64_t* bailed);
/* release the calling thread's staging workspace */
void ctd_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* CTD_B200_H */
