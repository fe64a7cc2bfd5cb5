// Autograd nodes of the two per-layer training ops in C++ (torch::autograd::Function): the sparse
// convolution (reference: ConvolutionFunction, torchsparse/nn/functional/conv.py:16-80) and the
// training-mode BatchNorm + shortcut + ReLU over sparse rows (nn.BatchNorm1d -> add -> relu,
// linkencoder.py:26-37, 64-91).  The arithmetic is liblinkb200's (the same C-ABI entry points the
// python Functions call: lk_conv_tc_pack_weights_ex, lk_conv_tc_fwd_plan, lk_conv_wgrad_tc,
// lk_bn_train_fwd / lk_bn_train_bwd) -- this file only moves the per-layer GLUE out of the
// interpreter: a training step of the encoder runs ~34 convs and ~30 BatchNorms forward and backward,
// the step is paced by the host, and the backward of a python autograd.Function re-enters the
// interpreter (GIL, argument marshalling through ctypes) for every node.  Here the backward nodes run
// on the autograd engine's thread without touching python.
//
// The library is NOT linked: python hands over the entry points' addresses (ctypes) once, so the
// extension always drives the liblinkb200.so the rest of the package loaded.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>

#include <stdexcept>
#include <string>

#include "../../include/linkb200.h"

namespace {

using pack_ex_t = int (*)(const float*, int, int, int, int, int, int, int, float*, lk_stream_t);
using conv_plan_fwd_t = int (*)(const float*, const float*, const int32_t*, const int32_t*, const uint32_t*, int64_t, int,
                                int, int, const lk_conv_epilogue_t*, float*, lk_stream_t);
using wgrad_t = int (*)(const float*, const float*, const int32_t*, const int32_t*, const uint32_t*, int64_t, int, int, int,
                        float*, int, lk_stream_t);
using bn_fwd_t = int (*)(const float*, const float*, int64_t, int, const float*, const float*, float, float, int, float*,
                         float*, int64_t*, float*, float*, float*, void*, int64_t, lk_stream_t);
using bn_bwd_t = int (*)(const float*, const float*, const float*, int64_t, int, const float*, const float*, const float*,
                         float*, float*, float*, float*, void*, int64_t, lk_stream_t);
using last_error_t = const char* (*)();

struct Fns {
  pack_ex_t pack_ex = nullptr;
  conv_plan_fwd_t conv_fwd = nullptr;
  wgrad_t wgrad = nullptr;
  bn_fwd_t bn_fwd = nullptr;
  bn_bwd_t bn_bwd = nullptr;
  last_error_t last_error = nullptr;
} g;

void bind(const std::string& name, int64_t addr) {
  void* p = reinterpret_cast<void*>(addr);
  if (name == "lk_conv_tc_pack_weights_ex") g.pack_ex = reinterpret_cast<pack_ex_t>(p);
  else if (name == "lk_conv_tc_fwd_plan") g.conv_fwd = reinterpret_cast<conv_plan_fwd_t>(p);
  else if (name == "lk_conv_wgrad_tc") g.wgrad = reinterpret_cast<wgrad_t>(p);
  else if (name == "lk_bn_train_fwd") g.bn_fwd = reinterpret_cast<bn_fwd_t>(p);
  else if (name == "lk_bn_train_bwd") g.bn_bwd = reinterpret_cast<bn_bwd_t>(p);
  else if (name == "lk_last_error") g.last_error = reinterpret_cast<last_error_t>(p);
  else throw std::runtime_error("link_b200 autograd extension: unknown entry point " + name);
}

bool ready() { return g.pack_ex && g.conv_fwd && g.wgrad && g.bn_fwd && g.bn_bwd && g.last_error; }

void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + " failed (code " + std::to_string(rc) + "): " + g.last_error());
}

lk_stream_t cur_stream() { return (lk_stream_t)c10::cuda::getCurrentCUDAStream().stream(); }

const float* fptr(const at::Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; }
const int32_t* iptr(const at::Tensor& t) { return t.defined() ? t.data_ptr<int32_t>() : nullptr; }
const uint32_t* uptr(const at::Tensor& t) { return t.defined() ? reinterpret_cast<const uint32_t*>(t.data_ptr<int32_t>()) : nullptr; }

// packed tensor-core image of w: layout 1 = [K, Cin, Cout] (the parameter), layout 0 = [K, Cout, Cin]
at::Tensor pack(const at::Tensor& w, int layout, bool flip) {
  const int k = (int)w.size(0);
  const int ci = (int)(layout ? w.size(1) : w.size(2)), co = (int)(layout ? w.size(2) : w.size(1));
  at::Tensor img = at::empty({(int64_t)k * 2 * ci * co}, w.options());
  check(g.pack_ex(w.data_ptr<float>(), k, ci, co, ci, co, layout, flip ? 1 : 0, img.data_ptr<float>(), cur_stream()),
        "lk_conv_tc_pack_weights_ex");
  return img;
}

at::Tensor run_conv(const at::Tensor& in, const at::Tensor& img, const at::Tensor& nbr, const at::Tensor& perm,
                    const at::Tensor& mask, int64_t n_out, int c_in, int c_out, int precision) {
  at::Tensor out = at::empty({n_out, (int64_t)c_out}, in.options());
  if (n_out == 0) return out;
  lk_conv_epilogue_t ep = {nullptr, nullptr, nullptr, 0, precision};
  check(g.conv_fwd(in.data_ptr<float>(), img.data_ptr<float>(), nbr.data_ptr<int32_t>(), iptr(perm), uptr(mask), n_out,
                   (int)nbr.size(0), c_in, c_out, &ep, out.data_ptr<float>(), cur_stream()),
        "lk_conv_tc_fwd_plan");
  return out;
}

// out[o] = sum_k feats[nbr[k, o]] @ weight[k]     (non-transposed conv on an output-stationary map)
//   dgrad: submanifold maps run on the FORWARD map (and plan) with the offsets of W reversed
//          (inv[k] == nbr[K-1-k]); other maps on the inverted map `inv` [K, n_in] (no plan)
//   wgrad: lk_conv_wgrad_tc over (wg_nbrp, wg_perm, wg_masks) = KernelMap.wgrad_relation(False)
struct ConvFunction : public torch::autograd::Function<ConvFunction> {
  // (optional tensors: an UNDEFINED at::Tensor argument of apply() is not accepted by the autograd wrapper)
  using OptT = c10::optional<at::Tensor>;
  static at::Tensor forward(torch::autograd::AutogradContext* ctx, at::Tensor feats, at::Tensor weight, at::Tensor nbr,
                            OptT perm_, OptT mask_, OptT inv_, OptT wg_nbrp_, OptT wg_perm_, OptT wg_masks_, bool subm,
                            int64_t precision, int64_t wgrad_slots) {
    auto o = [](const OptT& t) { return t.has_value() ? *t : at::Tensor(); };
    const at::Tensor perm = o(perm_), mask = o(mask_), inv = o(inv_), wg_nbrp = o(wg_nbrp_), wg_perm = o(wg_perm_),
                     wg_masks = o(wg_masks_);
    feats = feats.contiguous();
    weight = weight.contiguous();
    const int c_in = (int)weight.size(1), c_out = (int)weight.size(2);
    TORCH_CHECK(feats.size(1) == c_in, "Input feature size and kernel size mismatch");
    at::Tensor img = pack(weight, 1, false);
    at::Tensor out = run_conv(feats, img, nbr, perm, mask, nbr.size(1), c_in, c_out, (int)precision);
    ctx->save_for_backward({feats, weight, nbr, perm, mask, inv, wg_nbrp, wg_perm, wg_masks});
    ctx->saved_data["subm"] = subm;
    ctx->saved_data["precision"] = precision;
    ctx->saved_data["slots"] = wgrad_slots;
    return out;
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::variable_list grads) {
    auto sv = ctx->get_saved_variables();
    const at::Tensor &feats = sv[0], &weight = sv[1], &nbr = sv[2], &perm = sv[3], &mask = sv[4], &inv = sv[5],
                     &wg_nbrp = sv[6], &wg_perm = sv[7], &wg_masks = sv[8];
    const bool subm = ctx->saved_data["subm"].toBool();
    const int precision = (int)ctx->saved_data["precision"].toInt();
    const int slots = (int)ctx->saved_data["slots"].toInt();
    at::Tensor gout = grads[0].contiguous();
    if (gout.scalar_type() != at::kFloat) gout = gout.to(at::kFloat);
    const int k = (int)weight.size(0), c_in = (int)weight.size(1), c_out = (int)weight.size(2);
    at::Tensor gfeats, gweight;
    if (ctx->needs_input_grad(0)) {
      // dX = sum_k dY[to_in[k]] @ W[k]^T: the forward weight [K, Cin, Cout] IS the transposed operand
      // of a conv with Cout input channels and Cin output channels (layout 0)
      at::Tensor img = pack(weight, 0, subm);
      if (subm)
        gfeats = run_conv(gout, img, nbr, perm, mask, feats.size(0), c_out, c_in, precision);
      else
        gfeats = run_conv(gout, img, inv, at::Tensor(), at::Tensor(), feats.size(0), c_out, c_in, precision);
    }
    if (ctx->needs_input_grad(1)) {
      gweight = at::empty_like(weight);
      if (gout.size(0) > 0)
        check(g.wgrad(feats.data_ptr<float>(), gout.data_ptr<float>(), wg_nbrp.data_ptr<int32_t>(), iptr(wg_perm),
                      uptr(wg_masks), gout.size(0), k, c_in, c_out, gweight.data_ptr<float>(), slots, cur_stream()),
              "lk_conv_wgrad_tc");
      else
        gweight.zero_();
    }
    return {gfeats, gweight, at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor(),
            at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor()};
  }
};

// y = relu?( BN_train(x) [+ residual] ) with nn.BatchNorm1d's running-statistics updates
struct BatchNormActFunction : public torch::autograd::Function<BatchNormActFunction> {
  static at::Tensor forward(torch::autograd::AutogradContext* ctx, at::Tensor x, at::Tensor weight, at::Tensor bias,
                            c10::optional<at::Tensor> residual_, at::Tensor running_mean, at::Tensor running_var,
                            at::Tensor nbt, double eps, double momentum, bool relu) {
    const at::Tensor residual = residual_.has_value() ? *residual_ : at::Tensor();
    x = x.contiguous();
    const int64_t n = x.size(0);
    const int c = (int)x.size(1);
    at::Tensor res = residual.defined() ? residual.contiguous() : residual;
    at::Tensor y = at::empty_like(x);
    at::Tensor stats = at::empty({2, (int64_t)c}, x.options());
    at::Tensor ws = at::empty({2 * (int64_t)c}, x.options().dtype(at::kDouble));
    check(g.bn_fwd(x.data_ptr<float>(), fptr(res), n, c, fptr(weight), fptr(bias), (float)eps, (float)momentum, relu ? 1 : 0,
                   running_mean.defined() ? running_mean.data_ptr<float>() : nullptr,
                   running_var.defined() ? running_var.data_ptr<float>() : nullptr,
                   nbt.defined() ? nbt.data_ptr<int64_t>() : nullptr, stats.data_ptr<float>(),
                   stats.data_ptr<float>() + c, y.data_ptr<float>(), ws.data_ptr<double>(), 16 * (int64_t)c, cur_stream()),
          "lk_bn_train_fwd");
    ctx->save_for_backward({x, relu ? y : at::Tensor(), weight, stats});
    ctx->saved_data["has_res"] = residual.defined();
    return y;
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::variable_list grads) {
    auto sv = ctx->get_saved_variables();
    const at::Tensor &x = sv[0], &y = sv[1], &weight = sv[2], &stats = sv[3];
    const bool has_res = ctx->saved_data["has_res"].toBool();
    at::Tensor dy = grads[0].contiguous();
    if (dy.scalar_type() != at::kFloat) dy = dy.to(at::kFloat);
    const int64_t n = x.size(0);
    const int c = (int)x.size(1);
    at::Tensor dx = at::empty_like(x);
    at::Tensor dres = (has_res && ctx->needs_input_grad(3)) ? at::empty_like(x) : at::Tensor();
    at::Tensor dwb = at::empty({2, (int64_t)c}, x.options());
    at::Tensor ws = at::empty({2 * (int64_t)c}, x.options().dtype(at::kDouble));
    check(g.bn_bwd(dy.data_ptr<float>(), x.data_ptr<float>(), fptr(y), n, c, stats.data_ptr<float>(),
                   stats.data_ptr<float>() + c, fptr(weight), dx.data_ptr<float>(),
                   dres.defined() ? dres.data_ptr<float>() : nullptr, dwb.data_ptr<float>(), dwb.data_ptr<float>() + c,
                   ws.data_ptr<double>(), 16 * (int64_t)c, cur_stream()),
          "lk_bn_train_bwd");
    return {dx, dwb[0], dwb[1], dres, at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor()};
  }
};

at::Tensor conv(at::Tensor feats, at::Tensor weight, at::Tensor nbr, c10::optional<at::Tensor> perm,
                c10::optional<at::Tensor> mask, c10::optional<at::Tensor> inv, c10::optional<at::Tensor> wg_nbrp,
                c10::optional<at::Tensor> wg_perm, c10::optional<at::Tensor> wg_masks, bool subm, int64_t precision,
                int64_t wgrad_slots) {
  TORCH_CHECK(ready(), "link_b200 autograd extension: entry points not bound");
  return ConvFunction::apply(feats, weight, nbr, perm, mask, inv, wg_nbrp, wg_perm, wg_masks, subm, precision,
                             wgrad_slots);
}

at::Tensor batch_norm_act(at::Tensor x, at::Tensor weight, at::Tensor bias, c10::optional<at::Tensor> residual,
                          at::Tensor running_mean, at::Tensor running_var, at::Tensor nbt, double eps, double momentum,
                          bool relu) {
  TORCH_CHECK(ready(), "link_b200 autograd extension: entry points not bound");
  return BatchNormActFunction::apply(x, weight, bias, residual, running_mean, running_var, nbt, eps, momentum, relu);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("bind", &bind, "hand over the address of a liblinkb200 entry point");
  m.def("ready", &ready);
  m.def("conv", &conv);
  m.def("batch_norm_act", &batch_norm_act);
}
