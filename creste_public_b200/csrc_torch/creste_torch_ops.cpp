// creste_torch_ops.cpp -- the `creste::` dispatcher ops registered from C++ over the C ABI of libcreste_b200.so.
//
// The reference's deployment interface is a TorchScript file (scripts/runtime/compile.py:197-210:
// torch.jit.trace(...).save()) loaded by a C++ runtime.  creste_public_b200/torch_ops.py registers the ops for
// Python processes (that is what the tracer records); THIS library registers the same schemas without any Python:
// a libtorch program -- or a Python process that never imports creste_public_b200 -- loads it
// (torch::jit::load after dlopen / torch.ops.load_library) and runs the traced costmap model.
// Each op allocates its outputs with ATen on the inputs' device and launches on the current CUDA stream; the
// argument handling mirrors creste_public_b200/ops.py one to one.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/creste_b200.h"

namespace {

using at::Tensor;
using OptTensor = std::optional<Tensor>;

void check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, what, " failed (rc=", rc, "): ", creste_last_error());
}
const float* fp(const Tensor& t) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == at::kFloat, "creste ops need contiguous CUDA fp32 tensors");
  return t.data_ptr<float>();
}
const float* fpo(const OptTensor& t) { return (t.has_value() && t->defined()) ? fp(*t) : nullptr; }
void* stream() { return (void*)c10::cuda::getCurrentCUDAStream().stream(); }
Tensor f32(const Tensor& t) { return t.contiguous().to(at::kFloat); }

int act_id(const std::string& a) {
  static const std::map<std::string, int> m = {{"none", 0}, {"relu", 1}, {"swish", 2}, {"sigmoid", 3}};
  auto it = m.find(a);
  TORCH_CHECK(it != m.end(), "unknown activation ", a);
  return it->second;
}
int precision_id(const std::string& p) {
  static const std::map<std::string, int> m = {{"fp32", 0}, {"3xtf32", 1}, {"tf32", 2}, {"3xfp16", 4}, {"fp16", 5}};
  auto it = m.find(p);
  TORCH_CHECK(it != m.end(), "unknown precision ", p);
  return it->second;
}

Tensor conv2d(const Tensor& x_, const Tensor& w, int64_t K, int64_t R, int64_t S, int64_t stride, at::IntArrayRef pad,
              const OptTensor& scale, const OptTensor& shift, const OptTensor& gate, const OptTensor& residual,
              std::string act, bool out_nchw, std::string precision) {
  c10::cuda::CUDAGuard g(x_.device());
  Tensor x = x_.contiguous();
  const int N = x.size(0), H = x.size(1), W = x.size(2), C = x.size(3);
  const int P = (H + pad[0] + pad[1] - R) / stride + 1, Q = (W + pad[2] + pad[3] - S) / stride + 1;
  creste_conv_desc d = {N, H, W, C, (int)K, (int)R, (int)S, (int)stride, (int)pad[0], (int)pad[2], P, Q, act_id(act),
                        out_nchw ? 1 : 0, precision_id(precision)};
  Tensor out = at::empty(out_nchw ? std::vector<int64_t>{N, K, P, Q} : std::vector<int64_t>{N, P, Q, K}, x.options());
  const size_t n = creste_conv2d_workspace_bytes(&d);
  Tensor ws = at::empty({(int64_t)(n > 16 ? n : 16)}, x.options().dtype(at::kByte));
  check(creste_conv2d(&d, fp(x), fp(w), fpo(scale), fpo(shift), fpo(gate), fpo(residual), out.data_ptr<float>(),
                      ws.data_ptr(), n, stream()), "creste_conv2d");
  return out;
}

std::tuple<Tensor, Tensor> dwconv_bn_swish(const Tensor& x_, const Tensor& w, const Tensor& scale, const Tensor& shift,
                                           int64_t R, int64_t stride, at::IntArrayRef pad) {
  c10::cuda::CUDAGuard g(x_.device());
  Tensor x = x_.contiguous();
  const int N = x.size(0), H = x.size(1), W = x.size(2), C = x.size(3);
  const int P = (H + pad[0] + pad[1] - R) / stride + 1, Q = (W + pad[2] + pad[3] - R) / stride + 1;
  const int nparts = creste_dwconv_parts(N, P, Q, (int)R, (int)stride);   // same kernel choice as the Python mirror
  Tensor out = at::empty({N, P, Q, C}, x.options()), part = at::empty({N, nparts, C}, x.options());
  check(creste_dwconv_bn_swish(fp(x), fp(w), fp(scale), fp(shift), N, H, W, C, (int)R, (int)stride, (int)pad[0],
                               (int)pad[2], P, Q, out.data_ptr<float>(), part.data_ptr<float>(), nparts, stream()),
        "creste_dwconv_bn_swish");
  return {out, part};
}

Tensor se_gate(const Tensor& part, int64_t hw, const Tensor& w_red, const Tensor& b_red, const Tensor& w_exp,
               const Tensor& b_exp) {
  c10::cuda::CUDAGuard g(part.device());
  const int N = part.size(0), nparts = part.size(1), C = part.size(2), Csq = w_red.size(0);
  Tensor gate = at::empty({N, C}, part.options());
  check(creste_se_gate(fp(part), nparts, 1.0f / (float)hw, N, C, Csq, fp(w_red), fp(b_red), fp(w_exp), fp(b_exp),
                       gate.data_ptr<float>(), stream()), "creste_se_gate");
  return gate;
}

Tensor upsample_concat(const OptTensor& skip, const Tensor& x_, at::IntArrayRef out_hw, double ratio_h, double ratio_w,
                       bool x_first) {
  c10::cuda::CUDAGuard g(x_.device());
  Tensor x = x_.contiguous();
  const int N = x.size(0), Hi = x.size(1), Wi = x.size(2), Cx = x.size(3);
  const int Cs = (skip.has_value() && skip->defined()) ? (int)skip->size(3) : 0;
  Tensor out = at::empty({N, out_hw[0], out_hw[1], Cs + Cx}, x.options());
  check(creste_upsample_concat(fpo(skip), Cs, fp(x), N, Hi, Wi, Cx, (int)out_hw[0], (int)out_hw[1], (float)ratio_h,
                               (float)ratio_w, x_first ? 1 : 0, out.data_ptr<float>(), stream()), "creste_upsample_concat");
  return out;
}

std::tuple<Tensor, Tensor> maxpool2_concat(at::TensorList srcs, int64_t rows_out) {
  TORCH_CHECK(srcs.size() >= 1 && srcs.size() <= 3, "creste::maxpool2_concat: 1..3 sources");
  c10::cuda::CUDAGuard g(srcs[0].device());
  const int N = srcs[0].size(0), H = srcs[0].size(1), W = srcs[0].size(2);
  std::vector<Tensor> keep;
  const float* ptrs[3] = {nullptr, nullptr, nullptr};
  int chans[3] = {0, 0, 0}, Ct = 0;
  for (size_t i = 0; i < srcs.size(); ++i) {
    keep.push_back(srcs[i].contiguous());
    ptrs[i] = fp(keep.back());
    chans[i] = (int)keep.back().size(3);
    Ct += chans[i];
  }
  Tensor nhwc = at::empty({N, rows_out, W / 2, Ct}, srcs[0].options()), nchw = at::empty({N, Ct, rows_out, W / 2}, srcs[0].options());
  check(creste_maxpool2_concat(ptrs, chans, (int)srcs.size(), N, H, W, (int)rows_out, nhwc.data_ptr<float>(),
                               nchw.data_ptr<float>(), stream()), "creste_maxpool2_concat");
  return {nhwc, nchw};
}

Tensor nchw_to_nhwc(const Tensor& x_) {
  c10::cuda::CUDAGuard g(x_.device());
  Tensor x = x_.contiguous();
  Tensor out = at::empty({x.size(0), x.size(2), x.size(3), x.size(1)}, x.options());
  check(creste_nchw_to_nhwc(fp(x), x.size(0), x.size(1), x.size(2), x.size(3), out.data_ptr<float>(), stream()), "creste_nchw_to_nhwc");
  return out;
}

Tensor nhwc_to_nchw(const Tensor& x_) {
  c10::cuda::CUDAGuard g(x_.device());
  Tensor x = x_.contiguous();
  Tensor out = at::empty({x.size(0), x.size(3), x.size(1), x.size(2)}, x.options());
  check(creste_nhwc_to_nchw(fp(x), x.size(0), x.size(1), x.size(2), x.size(3), out.data_ptr<float>(), stream()), "creste_nhwc_to_nchw");
  return out;
}

std::tuple<Tensor, Tensor> depth_expectation(const Tensor& logits_, double dmin, double dmax, double out_div) {
  c10::cuda::CUDAGuard g(logits_.device());
  Tensor logits = logits_.contiguous();
  const int D = logits.size(-1);
  auto shp = logits.sizes().vec();
  shp.pop_back();
  Tensor metric = at::empty(shp, logits.options()), bins = at::empty(shp, logits.options().dtype(at::kLong));
  check(creste_depth_expectation(fp(logits), (int)(logits.numel() / D), D, (float)dmin, (float)dmax, (float)out_div,
                                 metric.data_ptr<float>(), bins.data_ptr<int64_t>(), stream()), "creste_depth_expectation");
  return {metric, bins};
}

std::tuple<Tensor, Tensor, Tensor> frustum_to_bev(const Tensor& depth_, const Tensor& p2p_, at::ArrayRef<double> pc_range,
                                                  at::ArrayRef<double> voxel) {
  c10::cuda::CUDAGuard g(depth_.device());
  Tensor depth = f32(depth_), p2p = f32(p2p_);
  const int N = depth.size(0), Hs = depth.size(1), Ws = depth.size(2), P = Hs * Ws;
  float rng[6], vox[2] = {(float)voxel[0], (float)voxel[1]};
  for (int i = 0; i < 6; ++i) rng[i] = (float)pc_range[i];
  Tensor xy = at::empty({N, P, 2}, depth.options()), z = at::empty({N, P}, depth.options()),
         mask = at::empty({N, P}, depth.options().dtype(at::kByte));
  check(creste_frustum_to_bev(fp(depth), fp(p2p), N, Hs, Ws, rng, vox, xy.data_ptr<float>(), z.data_ptr<float>(),
                              mask.data_ptr<uint8_t>(), stream()), "creste_frustum_to_bev");
  return {xy, z, mask};
}

Tensor zmlp_concat(const Tensor& feats_, const Tensor& z, const Tensor& w1, const Tensor& b1, const Tensor& w2,
                   const Tensor& b2) {
  c10::cuda::CUDAGuard g(feats_.device());
  Tensor feats = feats_.contiguous();
  const int C = feats.size(-1);
  auto shp = feats.sizes().vec();
  shp.back() = C + 32;
  Tensor out = at::empty(shp, feats.options());
  check(creste_zmlp_concat(fp(feats), fp(z.contiguous()), (int)(feats.numel() / C), C, fp(w1), fp(b1), fp(w2), fp(b2),
                           out.data_ptr<float>(), stream()), "creste_zmlp_concat");
  return out;
}

std::tuple<Tensor, Tensor, Tensor> splat_soft(const Tensor& xy_, const Tensor& feats_, const OptTensor& mask, int64_t H,
                                              int64_t W, double min_weight) {
  c10::cuda::CUDAGuard g(xy_.device());
  Tensor xy = f32(xy_), feats = f32(feats_);
  const int N = xy.size(0), P = xy.size(1), F = feats.size(-1);
  Tensor nhwc = at::empty({N, H, W, F}, xy.options()), nchw = at::empty({N, F, H, W}, xy.options()),
         dens = at::empty({N, 1, H, W}, xy.options());
  const size_t n = creste_splat_workspace_bytes(N, (int)H, (int)W, F);
  Tensor ws = at::empty({(int64_t)n}, xy.options().dtype(at::kByte));
  const uint8_t* m = (mask.has_value() && mask->defined()) ? mask->contiguous().data_ptr<uint8_t>() : nullptr;
  check(creste_splat_soft(fp(xy), fp(feats), m, N, P, F, (int)H, (int)W, (float)min_weight, nhwc.data_ptr<float>(),
                          nchw.data_ptr<float>(), dens.data_ptr<float>(), nullptr, ws.data_ptr(), n, stream()),
        "creste_splat_soft");
  return {nhwc, nchw, dens};
}

std::tuple<Tensor, Tensor, Tensor> proj_head(const Tensor& x_, const Tensor& w, const OptTensor& bias) {
  c10::cuda::CUDAGuard g(x_.device());
  Tensor x = x_.contiguous();
  const int N = x.size(0), H = x.size(1), W = x.size(2), C = x.size(3), K = w.size(0);
  Tensor pred = at::empty({N, H, W, K}, x.options()), pred_nchw = at::empty({N, K, H, W}, x.options()),
         x_nchw = at::empty({N, C, H, W}, x.options());
  check(creste_proj_head(fp(x), fp(w.contiguous()), fpo(bias), N, H, W, C, K, pred.data_ptr<float>(),
                         pred_nchw.data_ptr<float>(), x_nchw.data_ptr<float>(), stream()), "creste_proj_head");
  return {pred, pred_nchw, x_nchw};
}

}  // namespace

TORCH_LIBRARY(creste, m) {
  m.def("conv2d(Tensor x, Tensor w, SymInt K, SymInt R, SymInt S, SymInt stride, SymInt[] pad, Tensor? scale, Tensor? shift, "
        "Tensor? gate, Tensor? residual, str act, bool out_nchw, str precision) -> Tensor");
  m.def("dwconv_bn_swish(Tensor x, Tensor w, Tensor scale, Tensor shift, SymInt R, SymInt stride, SymInt[] pad) -> (Tensor, Tensor)");
  m.def("se_gate(Tensor chan_part, SymInt hw, Tensor w_red, Tensor b_red, Tensor w_exp, Tensor b_exp) -> Tensor");
  m.def("upsample_concat(Tensor? skip, Tensor x, SymInt[] out_hw, float ratio_h, float ratio_w, bool x_first) -> Tensor");
  m.def("maxpool2_concat(Tensor[] srcs, SymInt rows_out) -> (Tensor, Tensor)");
  m.def("nchw_to_nhwc(Tensor x) -> Tensor");
  m.def("nhwc_to_nchw(Tensor x) -> Tensor");
  m.def("depth_expectation(Tensor logits, float dmin, float dmax, float out_div) -> (Tensor, Tensor)");
  m.def("frustum_to_bev(Tensor depth, Tensor p2p, float[] pc_range, float[] voxel) -> (Tensor, Tensor, Tensor)");
  m.def("zmlp_concat(Tensor feats, Tensor z, Tensor w1, Tensor b1, Tensor w2, Tensor b2) -> Tensor");
  m.def("splat_soft(Tensor xy, Tensor feats, Tensor? mask, SymInt H, SymInt W, float min_weight) -> (Tensor, Tensor, Tensor)");
  m.def("proj_head(Tensor x, Tensor w, Tensor? bias) -> (Tensor, Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(creste, CUDA, m) {
  m.impl("conv2d", conv2d);
  m.impl("dwconv_bn_swish", dwconv_bn_swish);
  m.impl("se_gate", se_gate);
  m.impl("upsample_concat", upsample_concat);
  m.impl("maxpool2_concat", maxpool2_concat);
  m.impl("nchw_to_nhwc", nchw_to_nhwc);
  m.impl("nhwc_to_nchw", nhwc_to_nchw);
  m.impl("depth_expectation", depth_expectation);
  m.impl("frustum_to_bev", frustum_to_bev);
  m.impl("zmlp_concat", zmlp_concat);
  m.impl("splat_soft", splat_soft);
  m.impl("proj_head", proj_head);
}
