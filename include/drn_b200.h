/*
 * drn_b200.h -- C ABI of libdrn_b200.so: the B200 (sm_100a) kernels of the DRN-WSOD
 * per-image detection hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises or allocates;
 *   - return value 0 = OK, non-zero = error; drn_last_error() gives the thread-local message
 *     (the Python shim raises RuntimeError with it, mirroring the reference's assert/raise style);
 *   - activations are NHWC ("channels last"); dtype codes: DRN_F32 = 0, DRN_BF16 = 1;
 *   - "rows" are region proposals (R per image), "classes" K, refinement stages S.
 *
 * Each entry point cites the reference code it replaces (paths relative to the reference root;
 * WSL = projects/WSL/wsl/modeling).
 */
#ifndef DRN_B200_H
#define DRN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRN_F32 0
#define DRN_BF16 1
#define DRN_U8 2 /* images only (drn_resample_u8_fwd) */

typedef void* drn_stream_t;

int drn_version(void);
const char* drn_last_error(void);

/* First backbone conv fused with image normalisation.
 * Replaces WSL/meta_arch/rcnn.py:242-249 (preprocess_image: (img-mean)/std) +
 * WSL/backbone/resnet_ws.py:363-372,405-407 (stem.conv1 3x3 s2 + FrozenBN + ReLU) or
 * WSL/backbone/vgg.py:44-54,97-98 (plain1.conv1 3x3 s1 + bias + ReLU).
 * img: [3][H][W] fp32 (BGR 0..255) placed top-left on a zero Hp x Wp canvas (Hp>=H, Wp>=W: the
 * padding detectron2/structures/image_list.py:57-119 adds after normalisation);
 * w_packed: [3*3*3][Cout] fp32, row = (kh*3+kw)*3+cin; scale may be NULL;
 * out: [Ho][Wo][Cout] with Ho=(Hp+2-3)/stride+1; y = relu(conv*scale[c]+bias[c]). */
int drn_conv3x3_c3_fwd(const float* img_chw, int H, int W, int Hp, int Wp, const float* mean3_host,
                       const float* std3_host,
                       const float* w_packed, const float* scale, const float* bias, int Cout,
                       int stride, int relu, void* out_nhwc, int out_dtype, drn_stream_t stream);

/* Implicit-GEMM convolution / linear layer, fp32 SIMT (exact-fp32 mode).
 * Replaces detectron2/layers/wrappers.py:84-99 (Conv2d.forward = conv + norm + activation),
 * detectron2/layers/batch_norm.py:45-65 (FrozenBatchNorm2d as x*scale+bias), the residual add +
 * ReLU of WSL/backbone/resnet_ws.py:96-112,217-237 and nn.Linear + ReLU of
 * WSL/roi_heads/box_head.py:82-91 (a linear layer is the ksize=1, H=rows, W=1 case).
 * in: [N][H][W][Cin]; w: [ksize*ksize*Cin][Cout] (row = (kh*ksize+kw)*Cin+cin); stride 1,
 * padding = dilation*(ksize/2); scale may be NULL (=1); residual may be NULL; out: [N][H][W][Cout]
 * with row pitch ldo >= Cout elements.  Cin % 16 == 0, Cout % 64 == 0. */
int drn_conv_igemm_f32(const float* in, int N, int H, int W, int Cin, const float* w, int ksize,
                       int dilation, const float* scale, const float* bias, const float* residual,
                       int relu, float* out, int Cout, int ldo, drn_stream_t stream);

/* Tensor-core (tcgen05/TMEM/TMA) implicit-GEMM convolution / linear layer, bf16 in, fp32 accumulate.
 * Same reference lines as drn_conv_igemm_f32, plus the train-mode F.dropout of
 * WSL/roi_heads/box_head.py:90 fused into the epilogue (dropout_p > 0: same counter-based mask as
 * drn_dropout_inplace with the same seed, element index = row * Cout + col; effective seed =
 * dropout_seed + *dropout_seed_dev when the device pointer is non-NULL, so a captured CUDA graph
 * draws a fresh mask on every replay).
 * in: [N][H][W][Cin] bf16; w: [Cout][ksize*ksize*Cin] bf16 (K-major, K order (kh,kw,cin)); bias/scale
 * fp32 [Cout]; residual bf16 [N][H][W][Cout] (pitch Cout) or NULL; out dtype DRN_BF16 or DRN_F32 (F32:
 * no residual / dropout), row pitch ldo elements.  Cin % 64 == 0, Cout % 8 == 0, ldo % 8 == 0. */
int drn_conv_igemm_bf16_tc(const void* in, int N, int H, int W, int Cin, const void* w, int ksize,
                           int dilation, const float* scale, const float* bias, const void* residual,
                           int relu, void* out, int out_dtype, int Cout, int ldo, float dropout_p,
                           uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* workspace,
                           size_t workspace_bytes, drn_stream_t stream);
/* Scratch for the split-K schedules of deep-K GEMMs (fc6): drn_gemm_workspace_bytes() bytes, 16-byte aligned,
 * ZERO-INITIALISED once by the caller (the kernel resets the flags it uses), not shared by kernels that may run
 * concurrently.  NULL = whole tiles only.  drn_gemm_set_tail_split(1) (default 0; returns the previous
 * setting) lets a GEMM with K >= 16384 whose tiles end in a partial wave cut that wave's tiles into K-ranges
 * ("tail split-K": fp32 partials summed in a fixed order -- deterministic, but not bit-identical to the
 * one-piece accumulation).  Both schedules are measured negative results on fc6, kept for other shapes. */
size_t drn_gemm_workspace_bytes(void);
int drn_gemm_set_tail_split(int enabled);
/* Cap the number of SMs the persistent GEMM grids occupy (0 = all; returns the previous cap).  The grids are
 * one CTA per SM with a static tile schedule: when a collective (NCCL) holds some SMs for its whole duration,
 * a full-width GEMM grid would wait for them; a capped grid runs beside it. */
int drn_gemm_set_max_sms(int max_sms);

/* MaxPool2d(kernel 2, stride 1|2, padding 0), NHWC.
 * Replaces nn.MaxPool2d in WSL/backbone/resnet_ws.py:93-94,110-111,403,415 and vgg.py:93-94,108-109. */
int drn_maxpool2x2_nhwc(const void* in, int N, int H, int W, int C, int stride, int dtype, void* out,
                        drn_stream_t stream);

/* ROIPool(7x7, max) fused with the x(objectness+1) scaling.
 * Replaces detectron2/modeling/poolers.py:191-226 (ROIPooler.forward -> torchvision.ops.RoIPool)
 * and WSL/roi_heads/roi_heads_oicr.py:342-343 / roi_heads_wsddn.py:285-286.
 * feat: [h][w][C] (one image); boxes: [R][4] fp32 XYXY image px; objectness: [R] fp32 or NULL;
 * out: [R][49][C], bin-major ((ph*7+pw)*C + c) -- fc6's weight is permuted to that K order.
 * workspace: >= drn_roipool_workspace_bytes(h, w, C, dtype) bytes of scratch for the per-image
 * range-max tables (16-byte aligned), or NULL to use the direct per-cell scan (same results, slower
 * for large R).  Results are bit-identical either way (max has no rounding). */
size_t drn_roipool_workspace_bytes(int h, int w, int C, int dtype);
int drn_roipool_fwd(const void* feat_nhwc, int h, int w, int C, const float* boxes,
                    const float* objectness, int R, float spatial_scale, int dtype, void* out,
                    void* workspace, size_t workspace_bytes, drn_stream_t stream);

/* The two halves of drn_roipool_fwd, for callers that pool the proposals of one image in several row
 * blocks (so that the fc6 GEMM of block b overlaps the pooling of block b+1 on another stream):
 * drn_roipool_build_tables fills the per-image range-max tables once, drn_roipool_rows_fwd pools R rows
 * (boxes/objectness/out already offset to the block) from tables built earlier on an ordered stream;
 * max_ctas > 0 runs the gather as a persistent grid of that many CTAs (a small footprint per SM, so that
 * a concurrently running GEMM keeps its place), 0 = one CTA per (row, channel chunk).
 * drn_roipool_tables_supported(C, dtype) != 0 iff the channel count fits the table layout. */
int drn_roipool_tables_supported(int C, int dtype);
int drn_roipool_build_tables(const void* feat_nhwc, int h, int w, int C, int dtype, void* workspace,
                             size_t workspace_bytes, drn_stream_t stream);
int drn_roipool_rows_fwd(const void* feat_nhwc, int h, int w, int C, const float* boxes,
                         const float* objectness, int R, float spatial_scale, int dtype, void* out,
                         const void* tables, size_t tables_bytes, int max_ctas, drn_stream_t stream);

/* WSDDN dual-softmax MIL head + image-level BCE.
 * Replaces WSL/roi_heads/fast_rcnn.py:493-527 (softmax(cls,1)*softmax(det,0)), :689-700
 * (predict_probs_img: clamp(sum_r)), :317-329 (binary_cross_entropy_loss) for ONE image.
 * logits: [R][ld] fp32 with cls at columns [cls_off, cls_off+K), det at [det_off, det_off+K);
 * gt_onehot: [K] fp32; scores out [R][K]; img_score out [K]; loss out [1] =
 *   (mean_loss ? mean_k : sum_k) BCE(img_score, gt_onehot) * loss_scale  (loss_scale = 1/num_images).
 * row_ws: [2*R + K] fp32 scratch. */
int drn_wsddn_mil_fwd(const float* logits, int ld, int R, int K, int cls_off, int det_off,
                      const float* gt_onehot, int mean_loss, float loss_scale, float* scores,
                      float* img_score, float* loss, float* row_ws, drn_stream_t stream);

/* Pseudo-ground-truth mining of one OICR stage.
 * Replaces WSL/roi_heads/roi_heads_oicr.py:491-567 (get_pgt): for each image-level GT class g,
 * argmax over proposals of prev_scores[:, gt_classes[g]] (ties -> lowest index, NaN wins),
 * pgt box = that proposal's box (rederive != 0: passed through Box2BoxTransform.apply_deltas with the
 * given deltas (NULL = zeros), detectron2/modeling/box_regression.py:73-110, as stages k>=1 do via
 * fast_rcnn.py:1511-1532), weight = img_score[class].
 * prev_scores: [R][ld_prev]; gt_classes: [G] int64; deltas: row pitch ld_deltas, class c at columns
 * [4c,4c+4) (or [0,4) if cls_agnostic), or NULL; bbox_w: 4 floats (host).
 * Outputs: pgt_idx [G] int64, pgt_score [G], pgt_box [G][4], pgt_weight [G]. */
int drn_oicr_pgt(const float* prev_scores, int ld_prev, int R, const float* boxes,
                 const int64_t* gt_classes, int G, const float* img_score, int rederive,
                 const float* deltas, int ld_deltas, int cls_agnostic, const float* bbox_w_host,
                 int64_t* pgt_idx, float* pgt_score, float* pgt_box, float* pgt_weight,
                 drn_stream_t stream);

/* Proposal labelling against (pseudo) ground truth.
 * Replaces WSL/roi_heads/roi_heads.py:255-353 (label_and_sample_proposals; no sampling, :245-246),
 * detectron2/structures/boxes.py:329-361 (pairwise_iou), detectron2/modeling/matcher.py:61-103.
 * thresholds_host: nthr floats (ascending, without the -inf/+inf sentinels); labels_host: nthr+1
 * ints in {-1,0,1}.  Outputs: labels [R] int64 (class, K = background, -1 = ignore),
 * matched_idx [R] int64, counts [3] int32 = (#fg, #bg, #ignore).  G == 0 -> all background. */
int drn_label_proposals(const float* boxes, int R, const float* gt_boxes, const int64_t* gt_classes,
                        int G, int K, const float* thresholds_host, const int* labels_host, int nthr,
                        int64_t* labels, int64_t* matched_idx, int32_t* counts, drn_stream_t stream);

/* One OICR refinement stage: weighted softmax cross-entropy + next-stage probabilities.
 * Replaces WSL/roi_heads/roi_heads_oicr.py:381-397, fast_rcnn.py:1089-1096 (weights/valid),
 * :1128-1144 (weighted CE / #valid), :1098-1126 (_log_accuracy counters), :1561-1575 (softmax).
 * logits: [R][ld] with this stage's K+1 columns at col_off; labels/matched_idx from
 * drn_label_proposals; pgt_weight [G].  Outputs: probs [R][K+1]; loss [1]; stats [6] fp32 =
 * (#accurate, #fg, #fg_accurate, #false_negative, sum_r w*CE, #valid); weights [R] (proposal weights, may be NULL).
 * part_ws: [6*ceil(R/256)] fp32 scratch, counter: [1] uint32 zero-initialised (self-resetting). */
int drn_oicr_stage_fwd(const float* logits, int ld, int col_off, int R, int K, const int64_t* labels,
                       const int64_t* matched_idx, const float* pgt_weight, int G, float loss_scale,
                       float* probs, float* loss, float* stats, float* weights, float* part_ws,
                       uint32_t* counter, drn_stream_t stream);

/* Fused tail, part 1: drn_wsddn_mil_fwd + the stage-0 call of drn_oicr_pgt in ONE kernel (one CTA per class;
 * the class CTA that owns an image-level GT class also mines its pseudo GT: argmax over proposals of the
 * WSDDN score, box = that proposal, weight = image score).  Same arithmetic and reduction order as the
 * separate kernels (bit-identical outputs).  ws: [2*R + K] fp32; counter: [1] uint32 zero-initialised, self-resetting. */
int drn_wsddn_mil_pgt_fwd(const float* logits, int ld, int R, int K, int cls_off, int det_off,
                          const float* gt_onehot, int mean_loss, float loss_scale, const float* boxes,
                          const int64_t* gt_classes_img, int G, float* scores, float* img_score, float* loss,
                          int64_t* pgt_idx, float* pgt_score, float* pgt_box, float* pgt_weight, float* ws,
                          uint32_t* counter, drn_stream_t stream);

/* Fused tail, part 2: one refinement stage = drn_label_proposals (vs this stage's pseudo GT; optionally also the
 * first labelling vs the real GT, Gb >= 0) + drn_oicr_stage_fwd + the NEXT stage's drn_oicr_pgt (next_pgt_idx !=
 * NULL: argmax of this stage's probabilities, boxes re-derived through apply_deltas with `deltas` or zeros) in ONE
 * kernel; replaces the same reference lines.  counts / counts0: [3] int32 (#fg, #bg, #ignore).
 * part_ws: [(12 + 2 * G) * ceil(R/256)] 4-byte words of scratch. */
int drn_oicr_stage_fused_fwd(const float* logits, int ld, int col_off, int R, int K, const float* boxes,
                             const int64_t* gt_classes_img, int G, const float* pgt_box, const float* pgt_weight,
                             const float* thresholds_host, const int* labels_host, int nthr, float loss_scale,
                             const float* gt_boxes, const int64_t* gt_classes, int Gb, int64_t* labels0,
                             int64_t* matched0, int32_t* counts0, int64_t* labels, int64_t* matched_idx,
                             int32_t* counts, float* probs, float* loss, float* stats, float* weights,
                             const float* img_score, const float* deltas, int ld_deltas, int cls_agnostic,
                             const float* bbox_w_host, int64_t* next_pgt_idx, float* next_pgt_score,
                             float* next_pgt_box, float* next_pgt_weight, float* part_ws, uint32_t* counter,
                             drn_stream_t stream);

/* Fused tail, part 2 for ALL S refinement stages of one image in two launches (the default training path).  The pseudo GT of
 * stage k+1 is the per-class argmax of stage k's softmax (WSL/roi_heads/roi_heads_oicr.py:359-394, :491-567: get_pgt on the
 * detached predict_probs of the previous stage) -- a function of the LOGITS only -- so the S stages need not run one after
 * the other: launch 1 computes every stage's probabilities and the next stage's pseudo GT (grid.y = stage), launch 2 labels
 * the proposals against their stage's pseudo GT and reduces the weighted CE, #valid and the accuracy counters (grid.y = stage;
 * stage 0 also does the first labelling vs the real GT when Gb >= 0).  Same per-row arithmetic and reduction orders as S calls
 * of drn_oicr_stage_fused_fwd: bit-identical outputs.
 * Stage-major outputs: probs [S][R][K+1], labels / matched_idx / weights [S][R], counts [S][3], stats [S][6]; pgt_* [S][G(,4)]:
 * entry k = pseudo GT of stage k, entries 1.. are written, entry 0 is not touched (stage 0's comes in as pgt0_box / pgt0_weight
 * from drn_wsddn_mil_pgt_fwd).  Host arrays: col_offs / delta_offs (-1: no bbox_pred) / loss_cols [S] (stage k's loss goes to
 * loss[loss_cols[k]]), bbox_w [S][4] (row k = the weights that turn stage k's deltas into stage k+1's pseudo-GT boxes).
 * part_ws: [S * (12 + 2 * G) * ceil(R/256)] 4-byte words; counters: [2 * S] uint32 zero-initialised, self-resetting.
 * phases: 1 = launch 1 only, 2 = launch 2 only, 3 = both.  Launch 1 reads logits, boxes and gt_classes_img only (not img_score,
 * not pgt0_*): a caller may issue it on a second stream beside drn_wsddn_mil_pgt_fwd and join before launch 2. */
int drn_oicr_stages_fwd(const float* logits, int ld, int R, int K, int S, const int* col_offs, const int* delta_offs,
                        const float* bbox_w_host, const float* boxes, const int64_t* gt_classes_img, int G,
                        const float* img_score, int cls_agnostic, const float* pgt0_box, const float* pgt0_weight,
                        const float* thresholds_host, const int* labels_host, int nthr, float loss_scale,
                        const float* gt_boxes, const int64_t* gt_classes, int Gb, int64_t* labels0, int64_t* matched0,
                        int32_t* counts0, float* probs, int64_t* pgt_idx, float* pgt_score, float* pgt_box,
                        float* pgt_weight, int64_t* labels, int64_t* matched_idx, int32_t* counts, float* weights,
                        float* stats, float* loss, const int* loss_cols, float* part_ws, uint32_t* counters, int phases,
                        drn_stream_t stream);

/* Box regression loss of a refinement stage with REFINE_REG[k] (reg/ configs).
 * Replaces WSL/roi_heads/fast_rcnn.py:1146-1211 with smooth_l1(beta) and
 * detectron2/modeling/box_regression.py:38-71 (get_deltas).  deltas: [R][ld] at col_off, 4K wide
 * (or 4 if cls_agnostic); gt box of row r = pgt_box[matched_idx[r]].  loss = sum / R * loss_scale. */
int drn_oicr_boxreg_loss(const float* deltas, int ld, int col_off, int R, int K, int cls_agnostic,
                         const float* boxes, const float* pgt_box, const int64_t* labels,
                         const int64_t* matched_idx, const float* bbox_w_host, float beta,
                         float loss_scale, float* loss, float* part_ws, uint32_t* counter,
                         drn_stream_t stream);

/* Inference scores/boxes.
 * Replaces WSL/roi_heads/fast_rcnn.py:1577-1594 (predict_probs_K: mean of S softmaxes),
 * :1534-1559 (predict_boxes_K: mean deltas -> apply_deltas) and, for WSDDN heads, :668-687.
 * logits: [R][ld]; stage s has its K+1 class logits at col_offs_host[s]; delta_offs_host[s] >= 0
 * gives that stage's 4*nreg deltas (or -1 = zeros); nreg = K, or 1 if class-agnostic regression.
 * all_scores [R][K+1]; all_boxes [R][4*nreg]. */
int drn_oicr_infer(const float* logits, int ld, int R, int K, int nreg, int S, const int* col_offs_host,
                   const int* delta_offs_host, const float* boxes, const float* bbox_w_host,
                   float* all_scores, float* all_boxes, drn_stream_t stream);

/* Detections of one image from (all_scores, all_boxes): finite filter, drop the background column, clip
 * to the image, score threshold, per-class NMS, top-k -- all on the device with fixed-size outputs.
 * Replaces WSL/roi_heads/fast_rcnn.py:88-141 (fast_rcnn_inference_single_image),
 * detectron2/layers/nms.py:10-29 (batched_nms -> torchvision nms; the exact per-class form, see
 * csrc/drn_nms.cu for the restated arithmetic) and detectron2/structures/boxes.py (Boxes.clip).
 * all_scores: [R][K+1]; all_boxes: [R][4*nreg], nreg = K or 1 (class-agnostic); cap = number of output
 * slots (TEST.DETECTIONS_PER_IMAGE, or R*K for "all").  Outputs (first *num_out entries valid, sorted by
 * score descending, ties in (row, class) order): out_boxes [cap][4], out_scores [cap], out_classes [cap]
 * int64, out_rows [cap] int64 (index of the proposal row, the reference's filter_inds[:, 0]).
 * workspace: drn_detections_workspace_bytes(R, K) bytes, 16-byte aligned.  R <= 8192. */
size_t drn_detections_workspace_bytes(int R, int K);
int drn_detections_fwd(const float* all_scores, const float* all_boxes, int R, int K, int nreg, float img_h,
                       float img_w, float score_thresh, double nms_thresh, int cap, float* out_boxes,
                       float* out_scores, int64_t* out_classes, int64_t* out_rows, int32_t* num_out,
                       void* workspace, size_t workspace_bytes, drn_stream_t stream);

/* Train-mode dropout of the fc6/fc7 activations, in place: x = keep ? x/(1-p) : 0 with a
 * counter-based RNG keyed by (seed + *seed_dev (if non-NULL), element index).
 * Replaces F.dropout(p=0.5) in WSL/roi_heads/box_head.py:90 (mask not comparable to torch's RNG;
 * parity runs keep the box head in eval mode exactly like the survey's oracle). */
int drn_dropout_inplace(void* x, int64_t n, int dtype, float p, uint64_t seed, const uint64_t* seed_dev,
                        drn_stream_t stream);

/* ---- backward of the trainable tail ------------------------------------------------
 * The reference gets these from torch autograd; image scores, pseudo GT, proposal weights and next-stage
 * inputs are detached there (WSL/roi_heads/roi_heads_oicr.py:359-394), so every loss reaches only the logits
 * of its own head.  dlogits: [R][ld] fp32, same column layout as the forward's concatenated head logits;
 * grad_loss: device pointer to the upstream gradient of that loss (1 float) or NULL (= 1). */

/* d(loss_cls)/d(cls, det logits): backward of drn_wsddn_mil_fwd (fast_rcnn.py:493-527, :689-700, :317-329).
 * scores: [R][K] from the forward; g_ws: [K] fp32 scratch. */
int drn_wsddn_mil_bwd(const float* logits, int ld, int R, int K, int cls_off, int det_off, const float* scores,
                      const float* gt_onehot, int mean_loss, float loss_scale, const float* grad_loss,
                      float* dlogits, float* g_ws, drn_stream_t stream);

/* d(loss_cls_r{k})/d(cls_score_k logits): backward of drn_oicr_stage_fwd (fast_rcnn.py:1128-1144):
 * loss_scale * w_r (softmax - onehot) / *nvalid, zero for label -1.  probs [R][K+1], weights [R] and
 * nvalid (device, 1 float: stats[5] of the stage, summed over the images of the batch) come from the forward. */
int drn_oicr_stage_bwd(const float* probs, const int64_t* labels, const float* weights, const float* nvalid,
                       float loss_scale, const float* grad_loss, int R, int K, int ld, int col_off,
                       float* dlogits, drn_stream_t stream);

/* d(loss_box_reg_r{k})/d(bbox_pred_k deltas): backward of drn_oicr_boxreg_loss (fast_rcnn.py:1146-1211);
 * denom = total number of proposals in the batch.  Writes all 4K (or 4) delta columns (zeros off the fg class). */
int drn_oicr_boxreg_bwd(const float* deltas, int ld, int col_off, int R, int K, int cls_agnostic,
                        const float* boxes, const float* pgt_box, const int64_t* labels,
                        const int64_t* matched_idx, const float* bbox_w_host, float beta, float loss_scale,
                        float denom, const float* grad_loss, float* dlogits, drn_stream_t stream);

/* Gradient through ReLU + dropout, transposed for the weight-gradient GEMM:
 *   m[r][c] = grad[r][c] * (mask == NULL || mask[r][c] != 0 ? mul : 0)      (mask = the layer's forward output)
 *   out_t[c'][r] = m[r][c] for r < R, 0 for R <= r < ld_out;  c' = c, or (c % c49) * 49 + c / c49 when c49 > 0
 *   (bin-major pooled-feature columns -> the reference's (c, ph, pw) flatten order, box_head.py:87);
 *   out_masked[r][c] = m[r][c] if non-NULL (operand of the next input-gradient GEMM).
 * dtype codes per tensor; supported (grad, mask, out): (bf16,bf16,bf16) (f32,bf16,bf16) (bf16,bf16,f32) (f32,f32,f32). */
int drn_masked_transpose(const void* grad, int ld_grad, int grad_dtype, const void* mask, int ld_mask,
                         int mask_dtype, float mul, int R, int C, int c49, void* out_t, int ld_out,
                         void* out_masked, int ld_masked, int out_dtype, drn_stream_t stream);

/* out[row] = sum_i x[row][i], i < cols (bias gradients from a transposed gradient matrix; fixed order). */
int drn_rowsum(const void* x, int ld, int rows, int cols, int dtype, float* out, drn_stream_t stream);

/* [rows][49*c49] fp32, bin-major columns -> channel-major columns (fc6 weight gradient, exact-fp32 mode). */
int drn_permute_cols49(const float* in, float* out, int64_t rows, int c49, drn_stream_t stream);

/* One SGD update of a parameter tensor, fused with the refresh of its bf16 kernel-layout copy.
 * Arithmetic of torch.optim.SGD (the optimizer detectron2/solver/build.py:93-137 builds for the reference):
 *   d = grad + weight_decay * w;  buf = first_step ? d : momentum * buf + d;  d = nesterov ? d + momentum * buf : buf;
 *   w -= lr * d.
 * w/grad/momentum_buf: fp32 [rows][cols] in the parameter's layout, updated in place; packed_bf16 (may be NULL):
 * the [rows][cols] bf16 operand the GEMM kernels read -- same column order, or with c49 > 0 the fc6 permutation
 * (parameter columns (c, ph, pw) -> kernel columns (ph, pw, c), cols == 49 * c49, c49 % 64 == 0). */
int drn_sgd_step(float* w, const float* grad, float* momentum_buf, void* packed_bf16, int64_t rows, int64_t cols,
                 int c49, float lr, float momentum, float weight_decay, int nesterov, int first_step,
                 drn_stream_t stream);
/* ---- PCL refinement stage: projects/WSL/wsl/modeling/roi_heads/third_party/pcl.py:148-200
 * `_get_proposal_clusters` + wsl/layers/csrc/pcl_loss/pcl_loss_cpu.cpp:8-62 / :64-115 (the op behind wsl/layers/pcl_loss.py;
 * scaling :52,119) + the softmax of fast_rcnn.py PCLOutputs.predict_probs.  The cluster centres (k-means + IoU-graph cover,
 * third_party/pcl.py:62-145) are mined on the host, as in the reference, and passed in: center_boxes [P][4], center_classes [P]
 * (1-based: column of the refinement head, 0 = background), center_scores [P].
 * fwd: probs [R][K+1] = softmax(logits[:, col_off : col_off+K+1]); labels / assignment [R] int32 (0 / -1 below IoU 0.5),
 *   weights [R] (0 below IoU 0.1); per cluster pc_probs (mean clipped probability of its class over its members), pc_count,
 *   img_cls_loss_weights; loss[0] = loss_scale * (background term + cluster terms) / R.  terms_ws: P + 1 floats; counter: one
 *   zero-initialised uint32 (left at zero).
 * bwd: dlogits[:, col_off : col_off+K+1] = grad_loss[0] * d loss / d logits (overwrites those columns). */
int drn_pcl_stage_fwd(const float* logits, int ld, int R, int K, int col_off, const float* boxes, const float* center_boxes,
                      const int* center_classes, const float* center_scores, int P, float loss_scale, float* probs, int* labels,
                      float* weights, int* assignment, float* pc_probs, float* pc_count, float* img_cls_loss_weights, float* loss,
                      float* terms_ws, unsigned int* counter, drn_stream_t stream);
int drn_pcl_stage_bwd(const float* probs, int R, int K, const int* labels, const float* weights, const int* assignment,
                      const float* pc_probs, const float* pc_count, const float* img_cls_loss_weights, float loss_scale,
                      const float* grad_loss, int col_off, int ld, float* dlogits, drn_stream_t stream);

/* ---- data-parallel training of fc6.weight without an all-reduce (replaces the DistributedDataParallel
 * gradient all-reduce of detectron2/engine/defaults.py:279-282 + the replicated optimizer step of solver/build.py:93-137
 * for the one parameter that is 95 % of the gradient bytes) ----
 * drn_peer_get_handle: CUDA IPC handle (drn_peer_handle_bytes() bytes) of the cudaMalloc allocation `ptr` lives in, and
 *   ptr's byte offset inside it.  drn_peer_open / drn_peer_close: map / unmap another rank's allocation from such a
 *   handle; *base is the allocation's base address in this process (add the offset).
 * drn_gemm_bf16_tc_scatter: out[m][n] = sum_k a[m][k] * b[n][k] (+ bias[n]), bf16 operands, fp32 result, as
 *   drn_conv_igemm_bf16_tc (ksize 1, DRN_F32) -- but row m is owned by rank m / (M / n_peers) and every 128-row tile is
 *   stored by the GEMM's epilogue straight into the OWNER's window peer_windows[owner] (mapped peer memory, NVLink),
 *   slot my_rank: window layout [n_peers][M / n_peers][ldo] fp32.  M / n_peers % 128 == 0.  The weight gradient
 *   dW = dY^T X of a linear layer computed this way is reduce-scattered by the time the kernel (on every rank) has finished.
 * drn_sgd_step_sharded: the owner's part of the optimizer step: grad = (sum over the n_src slots, slot order) / n_src,
 *   torch.optim.SGD arithmetic (see drn_sgd_step) on rows [row0, row0 + rows) of the fp32 master w [*, cols] and on the
 *   momentum shard [rows][cols], then the refreshed bf16 kernel-layout rows (fc6 permutation, c49 > 0) stored into
 *   every rank's weight buffer packed_bf16[0 .. n_dst) (the all-gather of the update). */
int drn_peer_handle_bytes(void);
int drn_peer_get_handle(const void* ptr, void* ipc_handle, uint64_t* offset);
int drn_peer_open(const void* ipc_handle, void** base);
int drn_peer_close(void* base);
int drn_gemm_bf16_tc_scatter(const void* a, int M, int K, const void* b, int Nout, const float* bias,
                             void* const* peer_windows, int n_peers, int my_rank, int ldo, drn_stream_t stream);
int drn_sgd_step_sharded(float* w, float* momentum_shard, const float* slots, int n_src, void* const* packed_bf16,
                         int n_dst, int64_t row0, int64_t rows, int64_t cols, int c49, float lr, float momentum,
                         float weight_decay, int nesterov, int first_step, drn_stream_t stream);
/* The same bf16 kernel layout without an update (weight load / first use). */
int drn_pack_linear_bf16(const float* w, void* packed_bf16, int64_t rows, int64_t cols, int c49, drn_stream_t stream);

/* dtype / layout helpers used by the weight cache (not on the per-image path). */
int drn_cast_f32_to_bf16(const float* in, void* out, int64_t n, drn_stream_t stream);
int drn_cast_bf16_to_f32(const void* in, float* out, int64_t n, drn_stream_t stream);

/* ---- test-time augmentation (projects/WSL/wsl/modeling/test_time_augmentation_avg.py) ----
 *
 * drn_resample_u8_fwd: PIL Image.resize((new_w, new_h), BILINEAR) on an 8-bit image -- what
 * ResizeTransform.apply_image does for uint8 input (detectron2/data/transforms/transform.py:105-109, called by
 * DatasetMapperTTAAVG.__call__ test_time_augmentation_avg.py:121-124) -- fused with HFlipTransform.apply_image
 * (flip != 0: output column x <- new_w - 1 - x) and, for out_dtype DRN_F32, the uint8 -> float conversion of
 * GeneralizedRCNNWSL.preprocess_image.  Pillow's algorithm (src/libImaging/Resample.c): horizontal pass into a uint8
 * intermediate, then the vertical pass; out = clip8((2^21 + sum_k px[min + k] * coef[k]) >> 22).  A pass whose size
 * does not change is skipped.  Bit-exact (integer arithmetic).
 *   src_chw: uint8 [C][H][W] (planar);  out_chw: uint8 or fp32 [C][new_h][new_w];
 *   xbounds/ybounds: int32 [new][2] = (first source index, tap count);  xcoef/ycoef: int32 [new][ksize] fixed-point
 *   coefficients (host-built in double: precompute_coeffs + normalize_coeffs_8bpc), device pointers, may be NULL for
 *   a skipped pass;  tmp: uint8 [C][H][new_w] scratch, required when both passes run. */
int drn_resample_u8_fwd(const void* src_chw, int C, int H, int W, const int* xbounds, const int* xcoef, int xksize,
                        const int* ybounds, const int* ycoef, int yksize, int new_h, int new_w, void* tmp, int flip,
                        void* out_chw, int out_dtype, drn_stream_t stream);

/* drn_tta_accumulate: GeneralizedRCNNWithTTAAVG._get_augmented_boxes (test_time_augmentation_avg.py:286-309) for one
 * view: all_boxes [R][box_cols] go through the view's INVERSE transforms (fvcore Transform.apply_box: the four
 * corners through apply_coords in fp32 -- DRN_TTA_OP_RESIZE: x *= a, y *= b  (ResizeTransform.apply_coords,
 * transform.py:124-127, a = w / new_w, b = h / new_h rounded to fp32);  DRN_TTA_OP_HFLIP: x = a - x -- then min / max)
 * and are added to acc_boxes; all_scores [R][score_cols] are added to acc_scores.  view_index == 0 overwrites the
 * accumulators, view_index == n_views - 1 also divides by n_views (torch.mean over the views, summed in view order).
 * op_kind / op_a / op_b: HOST arrays of n_ops <= DRN_TTA_MAX_OPS entries, applied in order. */
#define DRN_TTA_MAX_OPS 4
#define DRN_TTA_OP_NOOP 0
#define DRN_TTA_OP_RESIZE 1
#define DRN_TTA_OP_HFLIP 2
int drn_tta_accumulate(const void* all_boxes, const void* all_scores, int R, int box_cols, int score_cols, int n_ops,
                       const int* op_kind, const float* op_a, const float* op_b, void* acc_boxes, void* acc_scores,
                       int view_index, int n_views, drn_stream_t stream);

/* ---- "fp32_tc" precision: fp32-accurate conv / linear layers on the bf16 tensor cores (csrc/drn_split.cu) ----
 * An fp32 value is the exact sum of three bf16 terms x1 + x2 + x3; sum_k x[k] w[k] is recovered to ~2^-24 from the six
 * leading bf16 x bf16 products.  tcgen05.mma truncates when it adds a K = 16 slice into its fp32 accumulator (measured
 * bias ~2^-25 per step, profiles/r1_fp32_tc_accumulator_truncation_v1.txt), so a layer of WSL/backbone/*.py or
 * roi_heads/box_head.py:82-91 runs as: ONE drn_conv_igemm_bf16_tc over the five correction products ("small" operand,
 * accumulator ~2^-8 of the result) + one drn_conv_igemm_bf16_tc per K-GROUP of the leading product x1 w1 ("big"
 * operand, <= ~64 MMA steps per accumulator), all with fp32 output, summed by drn_f32tc_reduce with round-to-nearest.
 *
 * drn_f32tc_split: x fp32 [rows][C] -> big bf16 [C / Cg][rows][Cg] (term x1, one dense matrix per K-group of Cg
 * channels) and small bf16 [rows][5 C] (planes x1 | x2 | x1 | x2 | x3, to meet weight planes w2 | w1 | w3 | w2 | w1).
 * drn_f32tc_reduce: y[r][c] = relu?(sum_{i < n_in} partials[i * part_stride + r * ld + c] + bias[c] + residual[r][c]),
 * fp32 adds in that order (bias / residual may be NULL).  C % 4 == 0. */
int drn_f32tc_split(const float* x, int64_t rows, int C, int Cg, void* big_bf16, void* small_bf16, drn_stream_t stream);
int drn_f32tc_reduce(const float* partials, int n_in, int64_t part_stride, int ld, const float* bias,
                     const float* residual, int relu, int64_t rows, int C, float* y, drn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
