/*
 * zsv_b200.h -- C ABI of libzsv_b200.so, the B200 (sm_100a) arithmetic under the reference's
 * nn.Module protocol for the R(2+1)D-18 / C3D training hot path and the zero-shot nearest-class search.
 *
 * The reference (damien911224/ZeroShotVideoClassification) has no FFI of its own: its seam is
 * PyTorch's nn.Module.forward / Tensor.backward (main.py:174, main.py:195).  Each entry point below
 * names the reference call site whose arithmetic it replaces.  Conventions:
 *   - every pointer is a CUDA device pointer owned by the caller (PyTorch); the library never
 *     allocates, frees or retains device memory and writes only through its out-parameters;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it, no
 *     host synchronisation, CUDA-graph capturable;
 *   - return value 0 = success, otherwise a zsv_status code; zsv_last_error() returns a
 *     thread-local message.  There is no CPU fallback: without a sm_100 device every compute
 *     entry point fails with ZSV_ERR_CUDA.
 *   - activations are bf16, channels-last NDHWC, channel pitch padded to a multiple of 8
 *     (zsv_cpad); pad lanes are written as zero by the producers.
 */
#ifndef ZSV_B200_H_
#define ZSV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum zsv_status {
    ZSV_OK = 0,
    ZSV_ERR_BAD_ARG = 1,     /* shape / stride / alignment not supported */
    ZSV_ERR_CUDA = 2,        /* CUDA runtime or driver call failed (message has the CUDA error) */
    ZSV_ERR_WORKSPACE = 3,   /* caller-provided workspace too small */
    ZSV_ERR_UNSUPPORTED = 4  /* configuration outside what the kernels implement */
} zsv_status;

/* Thread-local description of the last non-zero status returned on this thread. */
const char* zsv_last_error(void);
/* ABI version (bumped on any signature change). */
int zsv_abi_version(void);
/* Channel pitch used for a tensor with c channels: c rounded up to a multiple of 8. */
int zsv_cpad(int c);
/* Streaming multiprocessors of the current device (the persistent kernels launch one CTA per SM). */
int zsv_sm_count(void);
/* Number of CUDA kernels this library has launched in the calling process so far (monotonic). */
unsigned long long zsv_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Convolution descriptor.  Describes one nn.Conv3d of the reference:
 *   resnet.py:40-52 (Conv2Plus1D spatial 1x3x3 / temporal 3x1x1), resnet.py:181-184 (R2Plus1dStem),
 *   resnet.py:271 (1x1x1 stride-2 downsample), network.py:102-117 (C3D 3x3x3 + bias).
 * Strides must be 1 or 2 per dimension.
 * ---------------------------------------------------------------------------------------------- */
#define ZSV_CONV_X_NDHWC 0   /* input activation: bf16 [N][T][H][W][cpad(Cin)] */
#define ZSV_CONV_X_WFOLD 1   /* first-layer input: bf16 [N][T][H][W+8][8], 3..pw zero columns on the left
                                (produced by zsv_repack_input with layout=1); requires cpad(Cin)*kw <= 64 */
typedef struct zsv_conv_desc {
    int32_t N, T, H, W;      /* input extents (logical, unpadded) */
    int32_t Cin, Cout;       /* true channel counts */
    int32_t kt, kh, kw;      /* filter extents */
    int32_t st, sh, sw;      /* strides (1 or 2) */
    int32_t pt, ph, pw;      /* zero padding */
    int32_t x_layout;        /* ZSV_CONV_X_* */
} zsv_conv_desc;

/* Output extents of the convolution: out[0..2] = To, Ho, Wo. */
int zsv_conv3d_out_shape(const zsv_conv_desc* d, int32_t out[3]);

/* Bytes of the packed bf16 weight images.  which = 0: fprop image [tap][Cout][cpad(K)];
 * which = 1: dgrad image [tap][Cin][cpad(Cout)]. */
size_t zsv_conv3d_packed_weight_bytes(const zsv_conv_desc* d, int which);

/* fp32 [Cout][Cin][kt][kh][kw] (the state_dict layout, resnet.py:40) -> packed bf16 images.
 * Either output may be NULL. */
int zsv_conv3d_pack_weight(const zsv_conv_desc* d, const float* w, void* w_fprop, void* w_dgrad, void* stream);

/* The same for n convolutions in one launch (all weights of a network are re-packed after every optimizer step,
 * main.py:203): descs[n], w[n], w_fprop[n], w_dgrad[n] are HOST arrays of descriptors / device pointers; an output
 * pointer may be NULL. */
int zsv_conv3d_pack_weights(int n, const zsv_conv_desc* descs, const float* const* w, void* const* w_fprop,
                            void* const* w_dgrad, void* stream);

/* Inference (model.eval(), main.py:229-250): BatchNorm3d with running statistics is an affine map per output channel,
 * so it folds into the convolution: packed weights w'[co] = w[co]*gamma[co]/sqrt(running_var[co]+eps) and a bias
 * beta[co] - running_mean[co]*gamma[co]/sqrt(running_var[co]+eps), which zsv_conv3d_fprop applies (with ReLU and the
 * residual addend) in its epilogue: one kernel per conv -> BN -> (+shortcut) -> ReLU group.  One launch for n convs. */
typedef struct zsv_bn_fold {
    const float* gamma;         /* NULL = 1 */
    const float* beta;          /* NULL = 0 */
    const float* running_mean;
    const float* running_var;   /* NULL = do not fold this convolution */
    float* bias_out;            /* out: fp32 [Cout] */
    float eps;
} zsv_bn_fold;
int zsv_conv3d_pack_weights_folded(int n, const zsv_conv_desc* descs, const float* const* w, void* const* w_fprop,
                                   const zsv_bn_fold* fold, void* stream);

/* Number of rows of the BatchNorm partial-statistics buffers written by fprop: one per CTA of the persistent
 * kernel (each CTA accumulates its tiles in shared memory in a fixed order). */
int zsv_conv3d_stat_rows(const zsv_conv_desc* d);

/* Forward convolution (aten::conv3d reached from resnet.py:40-52,181-184,271; network.py:102-117).
 *   x       : bf16 input in d->x_layout
 *   w_fprop : packed weights (which = 0)
 *   y       : bf16 [N][To][Ho][Wo][cpad(Cout)]
 *   part_sum, part_sq : optional fp32 [stat_rows][cpad(Cout)] per-CTA sum / sum of squares of the
 *             bf16-rounded outputs (BatchNorm3d batch statistics, resnet.py:48); NULL to skip
 *   bias    : optional fp32 [Cout] (C3D, folded BatchNorm); addend: optional bf16 tensor shaped like y added before
 *             the activation (residual branch in the folded inference path, resnet.py:110); relu != 0 applies
 *             max(.,0) in the epilogue (network.py:147, resnet.py:111) */
int zsv_conv3d_fprop(const zsv_conv_desc* d, const void* x, const void* w_fprop, void* y, float* part_sum,
                     float* part_sq, const float* bias, const void* addend, int relu, void* stream);

/* Optional fusion of the first pass of BatchNorm3d backward into the data gradient.  The input of a convolution
 * is out = relu?(bn(y)) of the previous layer (resnet.py:48-49,95,182-186), so its data gradient g is exactly what
 * BatchNorm backward reduces over.  With this request dgrad writes dz = g * [y*scale+shift > 0] (or g when relu == 0)
 * instead of g, and every CTA leaves its per-channel sums of dz and dz*xhat in `partial`, so zsv_bn_bwd_finish can
 * go straight to the second pass; the separate reduction pass over g and y is gone. */
typedef struct zsv_bn_bwd_fuse {
    const void* y;          /* bf16 [N][T][H][W][cpad(Cin)]: pre-BatchNorm output of the layer that produced the input */
    const float* table;     /* fp32 [cpad(Cin)][4] = (scale, shift, invstd, -mean*invstd), from zsv_bn_finalize */
    int32_t relu;           /* != 0: that BatchNorm is followed by ReLU */
    float* partial;         /* out: fp32 [partial_rows][4][cpad(Cin)] (row r: [0] = sum dz, [1] = sum dz*xhat) */
    int32_t partial_rows;   /* capacity in rows; 8 * #SMs always suffices */
    int32_t rows_written;   /* out (host side): rows filled by this call */
} zsv_bn_bwd_fuse;

/* Data gradient (autograd of conv3d w.r.t. its input, triggered at main.py:195).
 *   dy : bf16 [N][To][Ho][Wo][cpad(Cout)];  w_dgrad : packed weights (which = 1)
 *   dx : bf16 [N][T][H][W][cpad(Cin)];  addend: optional bf16 tensor shaped like dx added in the
 *        epilogue (residual-branch gradient, resnet.py:103-111); bn_fuse: optional, see above (NULL = plain dgrad). */
int zsv_conv3d_dgrad(const zsv_conv_desc* d, const void* dy, const void* w_dgrad, void* dx, const void* addend,
                     zsv_bn_bwd_fuse* bn_fuse, void* stream);

/* Weight gradient (autograd of conv3d w.r.t. its weight).  dw is fp32 in the state_dict layout
 * [Cout][Cin][kt][kh][kw].  The workspace holds split-K partial tiles; zsv_conv3d_wgrad_workspace gives the
 * required size. */
size_t zsv_conv3d_wgrad_workspace(const zsv_conv_desc* d);
int zsv_conv3d_wgrad(const zsv_conv_desc* d, const void* x, const void* dy, float* dw, void* workspace,
                     size_t workspace_bytes, void* stream);
/* Bias gradient of a convolution (C3D, network.py:102-117): db[c] = sum over the rows of dy (bf16 [rows][cpad(C)]),
 * deterministic two-stage reduction through the workspace. */
size_t zsv_bias_grad_workspace(int C);
int zsv_bias_grad(const void* dy, float* db, long long rows, int C, void* workspace, size_t workspace_bytes,
                  void* stream);
/* ReLU backward on channels-last bf16: dz = g * [out > 0] (C3D conv+bias+ReLU, network.py:147-162).  bias_grad (optional,
 * fp32 [C]): the bias gradient sum over the rows of dz, taken in the same pass (workspace: zsv_bias_grad_workspace(C)). */
int zsv_relu_bwd(const void* g, const void* out, void* dz, long long rows, int C, float* bias_grad, void* workspace,
                 size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout conversion at the PyTorch boundary.
 * ---------------------------------------------------------------------------------------------- */
/* fp32 NCDHW clip batch (main.py:167, network.py:534-535) -> bf16 channels-last.
 * layout 0: [N][T][H][W][cpad(C)]; layout 1 (ZSV_CONV_X_WFOLD): [N][T][H][W+8][8] with `wpad_left`
 * zero columns on the left. */
int zsv_repack_input(const float* x, void* out, int N, int C, int T, int H, int W, int layout, int wpad_left,
                     void* stream);
/* Input pipeline stage fused with the layout conversion: decoded uint8 frames [N][T][Hs][Ws][3] (what the loader holds
 * before auxiliary/transforms.py:41-56) -> ToFloatTensorInZeroOne ((x/255-1)/2) -> Resize(resize_short) (bilinear,
 * align_corners=False, scale = resize_short/min(Hs,Ws)) -> crop x crop window at crop_ij[n] = (i, j) in the resized
 * frame -> optional horizontal flip (flip[n] != 0; NULL = none) -> bf16 W-folded layout [N][T][crop][crop+8][8] of
 * zsv_repack_input(layout=1).  crop_ij: device int32 [N][2]; flip: device uint8 [N].  A clip travels to the GPU as
 * uint8 (0.6 MB instead of 2.4 MB of fp32) and the CPU transform disappears. */
int zsv_clip_transform(const uint8_t* frames, void* out, int N, int T, int Hs, int Ws, int resize_short, int crop,
                       const int32_t* crop_ij, const uint8_t* flip, int wpad_left, void* stream);
/* bf16 NDHWC (pitch cpad(C)) <-> fp32 NCDHW, used by the per-module autograd shims and tests. */
int zsv_ndhwc_to_ncdhw(const void* x, float* out, int N, int C, int T, int H, int W, void* stream);
int zsv_ncdhw_to_ndhwc(const float* x, void* out, int N, int C, int T, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm3d (+ReLU, +residual), training mode (resnet.py:48,95,97,182,185,272; BasicBlock.forward
 * resnet.py:102-113).  rows = N*T*H*W positions, C true channels, pitch = cpad(C).
 * ---------------------------------------------------------------------------------------------- */
/* Reduce per-tile partials to batch mean / biased variance; emit scale = gamma*invstd,
 * shift = beta - mean*scale, mean, invstd (all fp32 [cpad(C)]); update running_mean / running_var
 * (momentum, unbiased variance) in place when they are non-NULL.  One launch: fp64 chunk sums, then the last block
 * of each 32-channel group (ticket counter) finishes.  The caller-provided workspace (zsv_bn_finalize_workspace
 * bytes) must have its first 1024 bytes ZERO before the first call; every call leaves them zero again, so calls
 * that follow each other on one stream may share the workspace.  bwd_table (optional): fp32 [cpad(C)][4] =
 * (scale, shift, invstd, -mean*invstd) per channel, the constants zsv_bn_bwd_fuse needs. */
size_t zsv_bn_finalize_workspace(int C);
int zsv_bn_finalize(const float* part_sum, const float* part_sq, int part_rows, int C, long long count,
                    const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                    float eps, float* scale, float* shift, float* mean, float* invstd, float* bwd_table,
                    void* workspace, size_t workspace_bytes, void* stream);
/* Eval mode: scale/shift from running statistics (main.py:229). */
int zsv_bn_eval_scale_shift(int C, const float* gamma, const float* beta, const float* running_mean,
                            const float* running_var, float eps, float* scale, float* shift, void* stream);
/* out = act( y*scale + shift  [+ y2*scale2 + shift2]  [+ residual] ), act = ReLU when relu != 0.
 * y2/scale2/shift2 (BatchNorm'd downsample branch, resnet.py:107-108) and residual (identity branch)
 * are optional. */
int zsv_bn_apply(const void* y, const float* scale, const float* shift, const void* y2, const float* scale2,
                 const float* shift2, const void* residual, void* out, long long rows, int C, int relu,
                 void* stream);
/* Backward of out = relu?(bn(y) [+ bn2(y2)] [+ residual]):
 *   g    : bf16 gradient w.r.t. out
 *   relu : 0 = no activation; 1 = ReLU, mask taken from the forward output `out`; 2 = ReLU of a single-branch unit,
 *          mask recomputed as y*mask_scale + mask_shift > 0 (the forward's scale/shift) so `out` is not re-read
 *   pass 1 (reduce): per-block partial sums of dz and dz*xhat for y (and y2) -> workspace
 *   pass 2 (apply) : dy = scale*(dz - mean(dz) - xhat*mean(dz*xhat)); optional dy2; optional dz
 *                    written out (gradient flowing into the identity residual).
 *   dgamma/dbeta are fp32 [C].  Workspace size from zsv_bn_bwd_workspace. */
size_t zsv_bn_bwd_workspace(int C);
int zsv_bn_bwd(const void* g, const void* out, int relu, const float* mask_scale, const float* mask_shift,
               const void* y, const float* mean, const float* invstd,
               const float* gamma, const void* y2, const float* mean2, const float* invstd2, const float* gamma2,
               void* dy, void* dy2, void* dz, float* dgamma, float* dbeta, float* dgamma2, float* dbeta2,
               long long rows, int C, void* workspace, size_t workspace_bytes, void* stream);

/* Second pass of BatchNorm backward when the sums came out of zsv_conv3d_dgrad (zsv_bn_bwd_fuse): reduces the
 * partial rows (fixed order, fp64) to dgamma / dbeta and writes dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)).
 * workspace: 4*cpad(C) floats. */
int zsv_bn_bwd_finish(const void* dz, const void* y, const float* mean, const float* invstd, const float* gamma,
                      const float* partial, int partial_rows, void* dy, float* dgamma, float* dbeta, long long rows,
                      int C, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embedding head: mean over (T,H,W) -> Linear(512,512) -> ReLU -> Linear(512,300) -> L2 normalise
 * (network.py:595-596, MLP network.py:603-618, F.normalize).  All fp32 except the bf16 feature map.
 * ---------------------------------------------------------------------------------------------- */
/* feat: bf16 [B][P][cpad(C)] (P = T*H*W positions).  Saves pooled [B][C], hidden [B][Hd], o-norm [B]. */
int zsv_head_fwd(const void* feat, int B, int P, int C, const float* w1, const float* b1, int Hd, const float* w2,
                 const float* b2, int E, float eps, float* pooled, float* hidden, float* onorm, float* emb,
                 void* stream);
/* Backward given demb [B][E]; grads for w1,b1,w2,b2 (fp32, overwritten) and dfeat (bf16 [B][P][cpad(C)]).
 * scratch: zsv_head_bwd_scratch(B, C, Hd, E) bytes. */
size_t zsv_head_bwd_scratch(int B, int C, int Hd, int E);
int zsv_head_bwd(const float* demb, const float* emb, const float* onorm, const float* pooled, const float* hidden,
                 int B, int P, int C, const float* w1, int Hd, const float* w2, int E, float eps, float* dw1,
                 float* db1, float* dw2, float* db2, void* dfeat, float* scratch, size_t scratch_bytes, void* stream);
/* Plain fp32 Linear (C3D fc6 / regressor, network.py:120,132,166,178): out = act(x W^T + b) with W [J][K] in the
 * state_dict layout.  Weight-streaming: every pass reads W once (per 24 batch rows); reductions that are split over
 * thread blocks are combined in a fixed order through the workspace (zsv_linear_workspace(B, K, J) bytes; forward
 * accepts NULL and then does not split).
 * Backward: `act` (optional) is the forward OUTPUT of a layer that ended in ReLU, its mask is applied to dy first;
 * dx, dw, db are each optional. */
size_t zsv_linear_workspace(int B, int K, int J);
int zsv_linear_fwd(const float* x, const float* w, const float* bias, float* out, int B, int K, int J, int relu,
                   void* workspace, size_t workspace_bytes, void* stream);
int zsv_linear_bwd(const float* dy, const float* x, const float* w, const float* act, int B, int K, int J, float* dx,
                   float* dw, float* db, void* workspace, size_t workspace_bytes, void* stream);
/* F.normalize(dim=-1) forward / backward (network.py:179, network.py:596). */
int zsv_l2norm_fwd(const float* o, float* emb, float* onorm, int B, int E, float eps, void* stream);
int zsv_l2norm_bwd(const float* demb, const float* emb, const float* onorm, float* dout, int B, int E, float eps,
                   void* stream);
/* MSELoss(mean) forward + gradient (main.py:130,179): loss[0] = mean((emb-target)^2),
 * demb = 2*(emb-target)/(B*E) * grad_scale. */
int zsv_mse_fwd_bwd(const float* emb, const float* target, int B, int E, float grad_scale, float* loss, float* demb,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step (main.py:131,200-203: torch.optim.Adam, no amsgrad), multi-tensor, CUDA-graph capturable.
 *   g' = g*grad_scale + weight_decay*p ; m += (1-beta1)*(g'-m) ; v = beta2*v + (1-beta2)*g'^2 ;
 *   p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps)
 * params / grads / exp_avg / exp_avg_sq / steps (/ numel) are HOST arrays of device pointers (sizes); steps[i] is a
 * DEVICE float holding t, the 1-based count of THIS update of tensor i (the caller increments it on the stream
 * beforehand).  lr_dev (optional) is a DEVICE float read at run time instead of `lr`, so that a learning-rate
 * schedule (main.py:133,374) reaches the replays of a captured graph.  grad_scale folds the 1/world of a summed
 * all-reduce or the 1/scale of a GradScaler (main.py:195-203) into the update; 1 = none.
 * ---------------------------------------------------------------------------------------------- */
typedef struct zsv_adam_hyper {
    float lr, beta1, beta2, eps, weight_decay, grad_scale;
    const float* lr_dev;
} zsv_adam_hyper;
int zsv_adam_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const long long* numel, const float* const* steps,
                  const zsv_adam_hyper* hyper, void* stream);
/* The same update for n convolution weights [Cout][Cin][kt][kh][kw] (descs[i]: only Cin, Cout, kt, kh, kw are read;
 * x_layout must be ZSV_CONV_X_NDHWC), FUSED with the bf16 re-pack the next forward / backward needs: the updated weights
 * are written as the fprop image w_fprop[i] and the dgrad image w_dgrad[i] (either may be NULL) of
 * zsv_conv3d_packed_weight_bytes -- the separate zsv_conv3d_pack_weights pass over the fp32 masters disappears. */
int zsv_adam_pack_step(int n, const zsv_conv_desc* descs, float* const* params, const float* const* grads,
                       float* const* exp_avg, float* const* exp_avg_sq, const float* const* steps,
                       void* const* w_fprop, void* const* w_dgrad, const zsv_adam_hyper* hyper, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Zero-shot nearest-class search (main.py:183, main.py:321-322: scipy cdist(...,'cosine') then
 * argmin / argsort[:, :k]).  fp32 inputs promoted to fp64 exactly like scipy; ties resolved towards
 * the lowest class index.  idx_out: int64 [N][k], k <= 8; dist_out (optional): fp64 [N][k].
 * ---------------------------------------------------------------------------------------------- */
int zsv_nearest_class(const float* emb, const float* cls, int N, int C, int D, int k, int64_t* idx_out,
                      double* dist_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MaxPool3d for C3D (network.py:103-118): kernel == stride, optional H/W padding with -inf.
 * argmax: one byte per output element (index of the selected position inside its window).
 * ---------------------------------------------------------------------------------------------- */
int zsv_maxpool3d_fwd(const void* x, void* y, uint8_t* argmax, int N, int T, int H, int W, int C, int kt, int kh,
                      int kw, int pt, int ph, int pw, void* stream);
/* relu_pooled (optional): the pooling OUTPUT y when the pooled tensor was a ReLU output (network.py:147-162): the
 * selected element equals y, so dy * [y > 0] is the ReLU backward of that element and the larger input is not re-read.
 * bias_grad (optional, fp32 [C]): sum over all positions of the gradient written to dx (the bias gradient of the
 * convolution in front of the ReLU), taken in the same pass; workspace: zsv_bias_grad_workspace(C) bytes. */
int zsv_maxpool3d_bwd(const void* dy, const uint8_t* argmax, const void* relu_pooled, void* dx, int N, int T, int H,
                      int W, int C, int kt, int kh, int kw, int pt, int ph, int pw, float* bias_grad, void* workspace,
                      size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ZSV_B200_H_ */
