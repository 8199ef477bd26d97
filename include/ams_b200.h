/* libams_b200 -- C ABI of the B200-native AMS student hot path.
 *
 * The reference (modelstreaming/ams) has no FFI: its boundary is `tf.Session.run` underneath the Python
 * class `SemanticNetwork`.  Each entry point below replaces one *role* of sess.run on that path; the
 * citation after each prototype is the reference call site (file:line in /root/reference) it stands in
 * for.  Plain pointers and sizes only; all `host` pointers are caller-owned host memory, device memory is
 * owned by the handle.  Every function returning int returns 0 on success; on failure a message is
 * available from ams_last_error() (thread-local).
 *
 * Threading: ams_enqueue() may be called from another thread concurrently with ams_train_step() /
 * ams_infer*() (reference: `_fill_queue` thread vs trainer thread, SemanticNetwork.py:230-231, :701 vs :260);
 * every other entry point assumes the caller serialises (reference `process_lock`, SemanticNetwork.py:70).
 */
#ifndef AMS_B200_H_
#define AMS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ams_net ams_net;

#define AMS_MAX_CLASSES 32
#define AMS_BN_MOVING 0  /* frozen client graph: inference-mode BN on moving statistics */
#define AMS_BN_BATCH 1   /* graph as shipped: FusedBatchNormV3(is_training=True), batch statistics */
#define AMS_FRAMES_U8 0
#define AMS_FRAMES_F32 1

typedef struct ams_config {
    int num_classes;                      /* logits channels of the graph: 19 (Cityscapes) or 21 (PASCAL VOC) */
    int graph_variant;                    /* 0 = deeplabv3_mobilenetv2_cityscapes, 1 = ..._pascalvoc2012 (BN decay table) */
    int height, width;                    /* frame size fed to the network (reference: width = 2*height, run.py:71) */
    int device;                           /* CUDA ordinal (reference: gpu_id -> visible_device_list, SemanticNetwork.py:74) */
    int class_count;                      /* number of selected classes (class_weights_exp == 1, SemanticNetwork.py:48-52) */
    int class_indices[AMS_MAX_CLASSES];   /* their logits channels, ascending (np.where(class_weights == 1)[0]) */
    int label_depth;                      /* one_hot depth for teacher labels: 19 (utils/graph_utils.py:15, :392) */
    int queue_capacity;                   /* staged input batches (reference FIFOQueue capacity 200); 0 = default 4 */
} ams_config;

const char* ams_last_error(void);
int ams_abi_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
long long ams_launch_count(void);

/* build: import graph + add heads/loss/Adam + create session -- SemanticNetwork.py:121-150, utils/graph_utils.py:338-533 */
ams_net* ams_create(const ams_config* cfg);
/* sess.close() -- SemanticNetwork.py:716-717 */
void ams_destroy(ams_net* net);
/* run all kernels of this handle on the caller's CUDA stream (cudaStream_t); NULL = the handle's own stream */
int ams_set_stream(ams_net* net, void* cuda_stream);
int ams_synchronize(ams_net* net);

/* ---- frozen hand-off (server -> client model).
 * get_frozen_graph + save_to_frozen_graph -- SemanticNetwork.py:706-714 -> trim_graph_frozen(kill_norms=True),
 * utils/graph_utils.py:79-126: writes the client model to `path`.  The reference writes a TF GraphDef with constants;
 * this library has no GraphDef, so the container ("AMSFRZ01") holds what that graph holds: every variable of the
 * student (272 tensors, tf.global_variables() order) to be run with inference-mode BatchNorm on the MOVING statistics. */
int ams_export_frozen(ams_net* net, const char* path);
/* frozen branch of SemanticNetwork.__init__ -- SemanticNetwork.py:80-118: a handle built from an exported client model.
 * cfg supplies height / width / device / selected classes / queue capacity; num_classes and graph_variant come from
 * the file (cfg->num_classes must be 0 or match).  The handle is inference-only: ams_train_* fail on it
 * ("Can't train frozen graph!!!", SemanticNetwork.py:217); run it with AMS_BN_MOVING. */
ams_net* ams_create_frozen(const char* path, const ams_config* cfg);
int ams_is_frozen(const ams_net* net);

/* ---- checkpoint variable layout: SaveHelper, utils/utils.py:10-49; names are TF variable names with ':0' */
int ams_num_tensors(const ams_net* net);
/* i-th variable in `tf.global_variables()` order of the shipped graph; shape4 zero-padded to 4 dims */
int ams_tensor_info(const ams_net* net, int index, char* name, int name_capacity, int shape4[4], int* ndim,
                    int* trainable, long long* arena_offset);
/* restore_vars feed of one variable -- utils/utils.py:44-47; KeyError semantics: unknown name fails */
int ams_set_tensor(ams_net* net, const char* name, const float* host, long long count);
/* save_vars fetch of one variable -- utils/utils.py:20-28; also serves '<var>/Adam:0', '<var>/Adam_1:0',
 * 'beta1_power:0', 'beta2_power:0' (the optimizer slots SemanticNetwork.get_vars returns) */
int ams_get_tensor(ams_net* net, const char* name, float* host, long long count);
/* whole trainable arena in tf.trainable_variables() order (2,113,043 floats for Cityscapes) in one copy */
long long ams_trainable_count(const ams_net* net);
int ams_get_trainable(ams_net* net, float* host);
int ams_set_trainable(ams_net* net, const float* host);
/* tf.global_variables_initializer for the Adam slots only (test hook; the reference never resets them) */
int ams_reset_optimizer(ams_net* net);

/* ---- input queue: sess.run(fill_input_buffer, feed) -- SemanticNetwork.py:176, :204, :701.
 * frames: [n,height,width,3] u8 or f32 (values 0..255); labels: [n,height,width] u8 teacher ids (ids outside the
 * selected classes, e.g. 255, are ignored pixels), NULL = all-ignored (predict_input enqueues zeros, :178).
 * Copies before returning; callable concurrently with ams_train_step / ams_infer*. */
int ams_enqueue(ams_net* net, const void* frames, int frames_dtype, const uint8_t* labels, int n);
/* Frame ingest on the device: the reference resizes every camera frame on the host before feeding it --
 * cv2.resize(frame, (W, H)) [INTER_LINEAR] + cv2.cvtColor(BGR2RGB) and, for teacher label maps,
 * cv2.resize(label, (W, H), interpolation=cv2.INTER_NEAREST)  (run.py:181-183, :263-264, :415-421).
 * frames: [n,src_h,src_w,3] u8 as decoded (BGR when bgr_to_rgb != 0); labels: [n,lab_h,lab_w] u8 or NULL.
 * Results are bit-identical to OpenCV's uint8 paths; same queue semantics and threading rules as ams_enqueue. */
int ams_enqueue_raw(ams_net* net, const uint8_t* frames, int src_h, int src_w, int bgr_to_rgb, const uint8_t* labels,
                    int lab_h, int lab_w, int n);
int ams_queue_size(ams_net* net);
/* drops every staged batch (returns how many): what a caller does after a failed training phase so that stale training
 * batches are not dequeued by the next predict_input (the reference would leave them in its FIFOQueue) */
int ams_queue_clear(ams_net* net);

/* ---- inference on the oldest queued batch.
 * sess.run(predictions) -- SemanticNetwork.py:173 (frozen), :179 (batch statistics).  out_labels int32 [n,H,W] */
int ams_infer(ams_net* net, int bn_mode, int32_t* out_labels);
/* reset_conf_mat + sess.run([predictions, update_op, loss]) -- SemanticNetwork.py:198-208.
 * out_confmat int64 [class_count*class_count], rows = teacher labels, cols = predictions; out_loss may be NaN */
int ams_infer_metric(ams_net* net, int bn_mode, int32_t* out_labels, int64_t* out_confmat, float* out_loss);
/* reset_conf_mat + sess.run(cross_update_op) -- SemanticNetwork.py:188-189 (ASR phi-score) */
int ams_confmat_labels(ams_net* net, const uint8_t* labels_before, const uint8_t* labels_after, long long n,
                       int64_t* out_confmat);

/* ---- distillation step on the oldest queued batch: sess.run({train_node, loss}) -- SemanticNetwork.py:260.
 * forward (batch-stat BN) + backward + BN moving-average update + TF1 Adam on every coordinate; when masked != 0
 * only coordinates with mask==1 are written back (backup -> minimize -> where(mask, new, backup),
 * utils/graph_utils.py:483-493).  out_loss = pre-update loss. */
int ams_train_step(ams_net* net, float lr, int masked, float* out_loss);
/* the same step without the host synchronisation: everything is enqueued on the handle's stream and the loss is copied
 * to out_loss_pinned (page-locked host memory, may be NULL) when the stream gets there -- the reference only prints the
 * loss (SemanticNetwork.py:261), so a training phase needs no per-step round trip; ams_synchronize() before reading */
int ams_train_step_async(ams_net* net, float lr, int masked, float* out_loss_pinned);
/* feed of the 164 mask placeholders (SemanticNetwork.py:255-257) as ONE byte map in trainable-arena order;
 * NULL = all ones */
int ams_set_mask(ams_net* net, const uint8_t* host_mask);
int ams_get_mask(ams_net* net, uint8_t* host_mask);
/* `_before = save_vars(...)` of get_train_mask -- SemanticNetwork.py:304: snapshot kept on the device */
int ams_snapshot_before(ams_net* net);
/* coord_desc_auto selection after iteration 0 -- SemanticNetwork.py:266-288: delta = |after - before|,
 * thr = np.percentile(delta, 100*(1-coord_frac)) (NumPy 1.19 semantics), mask = delta > thr, unselected
 * coordinates reverted to `before`; the mask becomes the current mask. */
int ams_select_topk(ams_net* net, double coord_frac, long long* out_kept, float* out_threshold);
/* model delta wire format -- run.py:316-328: per variable np.packbits(mask), then per variable
 * params[mask].astype(float16).  Returns the byte length (and fills out_buf when capacity suffices). */
int ams_pack_delta(ams_net* net, uint8_t* out_buf, long long capacity, long long* out_len);
/* Client side of the model stream (the reference only sizes the delta file, run.py:316-336, and
 * ships a whole frozen graph per update, :339, :401-411): applies a delta produced by ams_pack_delta to the resident
 * parameters -- coordinates whose mask bit is set take the fp16 value (widened to fp32), all others are untouched.
 * BatchNorm moving statistics are not part of the delta.  The handle's mask becomes the delta's mask. */
int ams_apply_delta(ams_net* net, const uint8_t* delta, long long length, long long* out_updated);

/* ---- data-parallel hooks (no reference counterpart: the reference is single-GPU) */
/* forward+backward only; gradients stay in the device gradient arena as sums over the LOCAL valid pixels
 * (not divided by n_valid) so that ranks can be summed; out_n_valid / out_loss_sum are the local terms */
int ams_train_forward_backward(ams_net* net, long long* out_n_valid, double* out_loss_sum);
/* asynchronous form of the exchange step: with both outputs NULL ams_train_forward_backward() does not synchronise and
 * leaves (n_valid, loss_sum) as two doubles on the device; the host plumbing allreduces that pair in place next to the
 * gradient arena, and ams_apply_optimizer_device() reads 1 / n_valid from it (0 valid pixels: scale 0, NaN loss).
 * out_loss_pinned (page-locked, may be NULL) receives the global mean loss when the stream gets there. */
void* ams_step_terms_device(ams_net* net);
int ams_apply_optimizer_device(ams_net* net, float lr, int masked, float* out_loss_pinned);
/* device pointer + length (floats) of the gradient arena, for an NCCL allreduce issued by the host plumbing */
void* ams_gradient_arena(ams_net* net, long long* count);
/* Bucketed gradient exchange overlapped with the backward pass ("1-4 buckets overlapped with backward;
 * late-layer buckets first").  The arena splits at ams_gradient_bucket_split() floats into an early-layer bucket
 * [0, split) and a late-layer bucket [split, count): every layer of the final-resolution stage + ASPP + logits (~90 % of
 * the coordinates), complete after ~45 % of the backward pass.  ams_gradient_bucket_wait() makes `cuda_stream` wait
 * until the late bucket of the most recently enqueued ams_train_forward_backward() is complete (an event recorded in
 * the middle of the step -- an external event node when the step is replayed as a CUDA graph), so the host plumbing
 * can run that bucket's allreduce on its own stream while the high-resolution layers are still in backward. */
long long ams_gradient_bucket_split(const ams_net* net);
int ams_gradient_bucket_wait(ams_net* net, void* cuda_stream);
/* Adam + mask with gradients scaled by grad_scale (= 1 / global n_valid) */
int ams_apply_optimizer(ams_net* net, float lr, int masked, float grad_scale);

/* Global-batch BatchNorm for the data-parallel step.  The reference computes FusedBatchNormV3(is_training=True)
 * statistics over the whole batch in one process (model.meta; utils/graph_utils.py:457), so a job that splits the
 * batch over GPUs must sum (sum z, sum z^2) forward and (sum g, sum g*z) backward over the ranks for each of the 54
 * BatchNorm layers.  Here the BN finalize kernels do that themselves: each rank pushes its fp64 pairs into every
 * peer's receive buffer over NVLink (CUDA IPC peer memory, 8-byte words carrying payload + epoch flag) and adds the
 * world's pairs in rank order -- no host or NCCL call inside the step, bit-identical statistics on every rank.
 *   ams_syncbn_init    allocates this rank's receive buffer and returns its 64-byte cudaIpcMemHandle_t
 *   ams_syncbn_connect takes the handles of all ranks (rank order, 64 bytes each; exchanged by the host plumbing),
 *                      maps the peers and switches the exchange on for ams_train_* (never for ams_infer*)
 *   ams_syncbn_enable  switches it off / on again (e.g. around a rank-local profiling step)
 *   ams_syncbn_status  synchronises; error != 0 means a peer did not arrive within the timeout (10 s, env
 *                      AMS_SYNCBN_TIMEOUT_MS) -- the kernels never hang, the step's results are then invalid.
 * All ranks must run the same sequence of training steps with the same per-rank batch size. */
int ams_syncbn_init(ams_net* net, int world, int rank, void* out_ipc_handle, int handle_capacity);
int ams_syncbn_connect(ams_net* net, const void* all_handles, int count);
int ams_syncbn_enable(ams_net* net, int on);
int ams_syncbn_status(ams_net* net, unsigned int* out_epoch, unsigned int* out_error);

/* Frozen inference runs the inverted-residual blocks (stride 1 and 2) whose project conv has <= 256 output channels (15 of the 17)
 * as ONE kernel each (expand GEMM with the expanded channels on the TMEM lanes -> BN/ReLU6 and the depthwise 3x3 in
 * registers, one channel per thread -> MN-major shared-memory A operand -> project GEMM): the 6C-wide tensors never reach
 * HBM and ams_get_activation() has nothing to return for the expand / depthwise layers of those blocks.  Same fp16
 * rounding points as the one-kernel-per-layer schedule; the two agree up to the fp32 accumulation order
 * (tests/test_net_gpu.py: argmax agreement > 99.5 %, logits rel-L2 < 3e-3; measured 1.00000 / 2e-4).  Default ON (env
 * AMS_BLOCK_FUSION=0: off).  ams_set_block_fusion(net, 0) restores the per-layer schedule, e.g. to inspect activations. */
int ams_set_block_fusion(ams_net* net, int on);

/* Frozen inference (AMS_BN_MOVING) of an even batch of >= 4 frames runs as two half batches on two streams inside one
 * CUDA graph: the 1/16-resolution layers launch fewer CTAs than there are SMs and every layer ends in a partial wave,
 * so two independent chains fill each other's gaps.  The half-batch plans are views of the full plan's buffers, so
 * ams_get_logits / ams_get_activation see the same data; predictions, confusion matrix and loss are bit-identical to
 * the unsplit run (tests/test_net_gpu.py).  Default ON (env AMS_NO_INFER_SPLIT=1: off). */
int ams_set_infer_split(ams_net* net, int on);

/* ---- teacher network (config C5): DeepLabv3+ / Xception-65, inference only -- the network behind
 * `sess.run(teacher['predictions'], feed_dict={teacher['images']: frame})` of extract_labels.py:84 (graph imported by
 * create_teacher, utils/graph_utils.py:129-152).  The teacher's .meta and weights are not in the reference repository;
 * the topology restates the public model-zoo definition (ams_b200/csrc/teacher.cu) and variables are addressed by the
 * TF names of that definition, so a checkpoint dict `{'<name>:0': ndarray}` loads with ams_teacher_set_tensor. */
typedef struct ams_teacher ams_teacher;
ams_teacher* ams_teacher_create(int num_classes, int device);
void ams_teacher_destroy(ams_teacher* t);
int ams_teacher_num_tensors(const ams_teacher* t);
int ams_teacher_tensor_info(const ams_teacher* t, int index, char* name, int name_capacity, int shape4[4], int* ndim);
int ams_teacher_set_tensor(ams_teacher* t, const char* name, const float* host, long long count);
/* frames [n,height,width,3] u8 RGB (the caller applies the reference's 1-px symmetric top/left pad, extract_labels.py:83);
 * out_labels int32 [n,height,width] = predictions:0; out_logits (may be NULL) fp32 [n, ceil(h/4), ceil(w/4), num_classes]
 * = logits/semantic/BiasAdd:0 before the final upsample.  Plans are rebuilt when the frame size changes. */
int ams_teacher_predict(ams_teacher* t, const uint8_t* frames, int n, int height, int width, int32_t* out_labels, float* out_logits);
/* host-only (no device): the variable table of the restated teacher graph, for checkpoint tooling and tests */
int ams_teacher_layout_num_tensors(int num_classes);
int ams_teacher_layout_tensor_info(int num_classes, int index, char* name, int name_capacity, int shape4[4], int* ndim);
/* device time (CUDA events on the teacher's stream) of one forward pass at the last predicted size, averaged over reps */
int ams_teacher_time_forward(ams_teacher* t, int reps, float* out_ms);

/* ---- parity hooks */
int ams_get_logits(ams_net* net, float* host, long long count);       /* low-res `semantic` [n,h,w,num_classes] of the last run */
int ams_get_gradients(ams_net* net, float* host);                     /* trainable-arena order, last train step */
int ams_num_layers(const ams_net* net);
int ams_layer_info(const ams_net* net, int index, char* name, int name_capacity, int* kind, int* cin, int* cout,
                   int* stride, int* dilation, int* act, float* bn_eps, float* bn_one_minus_decay, int* residual_from);
/* tensor of conv layer `index` from the last run as raw 16-bit words: which = 0 post-BN/act output (IEEE fp16),
 * 1 raw conv output (training; fp16), 2 gradient wrt the layer output (bf16) */
int ams_get_activation(ams_net* net, int index, int which, uint16_t* host_bits16, long long count);

/* ---- per-kernel-group device timing (CUDA events on the launching stream), for bench.py's roofline line.
 * report: one line per group "<tag> <launches> <total_ms> <algorithmic_bytes>"; returns the text length */
int ams_profile_enable(ams_net* net, int on);
int ams_profile_report(ams_net* net, char* buf, int capacity);

/* ---- host-only layout queries (no device needed): the variable / layer tables a handle with this graph would
 * have.  Lets a maintainer (and tests/test_layout.py) check the restated topology against model.meta. */
int ams_layout_num_tensors(int num_classes, int graph_variant);
int ams_layout_tensor_info(int num_classes, int graph_variant, int index, char* name, int name_capacity, int shape4[4],
                           int* ndim, int* trainable, long long* arena_offset);
int ams_layout_num_layers(int num_classes, int graph_variant);
int ams_layout_layer_info(int num_classes, int graph_variant, int index, char* name, int name_capacity, int* kind, int* cin,
                          int* cout, int* stride, int* dilation, int* act, float* bn_eps, float* bn_one_minus_decay,
                          int* residual_from);

/* ---- host-only diagnostics: the shared-memory tile the depthwise kernels pick for a geometry (no CUDA calls).
 * out8 = {tile_h, tile_w, tiles_x, tiles_y, channel_chunk, smem_bytes, ctas, threads (fwd) | strips per row (bwd)} */
int ams_debug_dw_tile(int n, int h, int w, int c, int ho, int wo, int stride, int dilation, int* out8);
int ams_debug_dw_bwd_tile(int n, int h, int w, int c, int ho, int wo, int stride, int dilation, int* out8);

/* ---- op-level entry points on raw DEVICE pointers (kernel unit tests; see tests/test_ops_gpu.py).
 * Storage types: forward activations and the 1x1 weight operands are IEEE fp16 (`_f16`), activation gradients bf16
 * (`_bf16`); accumulation, statistics, parameters and parameter gradients are fp32. */
/* grad_types == 0: forward conv (a, w fp16; out / residual fp16); != 0: data gradient (a, w bf16; out / residual bf16).
 * w_lo_16 != NULL: split weights, out = a*w^T + a*w_lo^T in one fp32 accumulator (w = fp16(W), w_lo = fp16(W - w)) */
int ams_op_conv1x1(const void* a_16, const void* w_16 /*[N][K]*/, int M, int N, int K, const float* scale,
                   const float* shift, const float* rowbias, int rows_per_image, const void* residual_16, int act,
                   void* out, int out_fp32, int ldc, int grad_types, const void* w_lo_16, void* stream);
/* one inverted-residual block of the FROZEN graph in a single kernel (ams_b200/csrc/fused_block.cu): stride 1 with
 * dilation 1 | 2, or stride 2 with dilation 1 (output [n,ceil(h/2),ceil(w/2),cout], TensorFlow 'SAME' padding, no skip):
 * out = BN3(project(relu6(BN2(depthwise3x3_dil(relu6(BN1(expand(x)))))))) [+ x]; x [n,h,w,cin] fp16; we [cexp][cin] and
 * wp [cout][cexp] fp16 (+ optional low planes of the split weights); wd [3,3,cexp] fp32; s1..s3 / t1..t3 folded BN scale / shift */
int ams_op_fused_block(const void* x_f16, int n, int h, int w_, int cin, int cexp, int cout, int dilation, int stride, const void* we_f16,
                       const void* we_lo_f16, const float* s1, const float* t1, const float* wd, const float* s2, const float* t2,
                       const void* wp_f16, const void* wp_lo_f16, const float* s3, const float* t3, int residual, void* out_f16,
                       void* stream);
/* diagnostics: device buffer (3 x 64 x 4 u64) that the next ams_op_fused_block fills with the timeline (ns) of CTA 0 */
int ams_debug_fused_timeline(void* device_buffer);
int ams_op_wgrad(const void* x_f16, int cin, const void* dz_bf16, int cout, long long M, float* dw, void* stream);
int ams_op_depthwise(const void* in_f16, const float* w, int n, int h, int w_, int c, int stride, int dilation,
                     const float* scale, const float* shift, int act, void* out_f16, void* stream);
/* cv2.resize on DEVICE buffers: [n,src_h,src_w,channels] u8 -> [n,dst_h,dst_w,channels]; nearest != 0: INTER_NEAREST
 * (1 channel), else INTER_LINEAR (1 or 3 channels; swap_rb exchanges channels 0 and 2 on the way out) */
int ams_op_resize_u8(const void* src, int n, int src_h, int src_w, int channels, void* dst, int dst_h, int dst_w, int nearest,
                     int swap_rb, void* stream);
/* tiled depthwise with the producer's BN+act applied while staging the input (in_scale/in_shift may be NULL) and the
 * batch statistics of the stored output: stats_out[2][c] fp64 DEVICE = column sums (sum, sum of squares) */
int ams_op_depthwise_fused(const void* in_f16, const float* w, int n, int h, int w_, int c, int stride, int dilation,
                           const float* in_scale, const float* in_shift, int in_act, void* out_f16, double* stats_out,
                           void* stream);
/* fused depthwise backward (see ams_b200/csrc/dw_tiled.cu): coef = [3][c] BN-backward coefficients of the depthwise
 * layer; zin = raw output of the producer whose BN (in_scale/in_shift) + activation is recomputed; bn_sums [2][c] fp64
 * DEVICE = column sums (masked gradient, masked gradient * zin) for the producer's BN backward */
int ams_op_depthwise_bwd_fused(const void* g_bf16, const void* z_f16, const float* scale, const float* shift, int act,
                               const float* coef, const void* zin_f16, const float* in_scale, const float* in_shift,
                               int in_act, const float* w, int n, int h, int w_, int c, int stride, int dilation,
                               void* gout_bf16, float* dw, double* bn_sums, void* stream);
int ams_op_depthwise_bwd(const void* x_f16, const void* dz_bf16, const float* w, int n, int h, int w_, int c, int stride,
                         int dilation, void* dx_bf16, float* dw, void* stream);
int ams_op_stem(const void* frames, int frames_dtype, int n, int h, int w_, const float* w, const float* scale,
                const float* shift, void* out_f16, void* stream);
int ams_op_stem_bwd(const void* frames, int frames_dtype, int n, int h, int w_, const void* dz_bf16, float* dw, void* stream);
int ams_op_bn_train(const void* z_f16, long long M, int C, const float* gamma, const float* beta, float eps, int act,
                    const void* residual_f16, void* y_f16, float* mean, float* rstd, void* stream);
int ams_op_bn_backward(const void* dy_bf16, const void* z_f16, long long M, int C, const float* gamma, const float* beta,
                       float eps, int act, void* dz_bf16, float* dgamma, float* dbeta, void* stream);
int ams_op_head_infer(const float* logits, int n, int h, int w_, int ldl, int H, int W, int class_count,
                      const int* class_indices, int label_depth, const uint8_t* labels, int32_t* pred, int64_t* confmat,
                      double* loss_sum, long long* n_valid, void* stream);
int ams_op_head_backward(const float* logits, int n, int h, int w_, int H, int W, int class_count, const int* class_indices,
                         int label_depth, const uint8_t* labels, float* dlogits /*[n,h,w,32]*/, float* loss, void* stream);
int ams_op_select(float* after, const float* before, long long n, double coord_frac, uint8_t* mask, long long* kept,
                  float* threshold, void* stream);
int ams_op_adam(float* p, const float* g, float* m, float* v, const uint8_t* mask, long long n, float lr, float beta1_power,
                float beta2_power, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMS_B200_H_ */
