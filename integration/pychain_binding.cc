// pychain_binding.cc -- the binding a maintainer of the reference's native module would add.
//
// Same module surface as /root/reference/pytorch_binding/src/pychain.cc:131-135
// (forward_backward, forward_backward_log_domain, set_verbose_level; same positional arguments, same return
// triple), but every CUDA call is forwarded to the C ABI of include/chain_b200.h instead of the reference's
// ChainComputation / ChainLogDomainComputation drivers.  This file is what INTEGRATION.md section 3 describes;
// tests/test_gpu_dropin.py compiles it with torch.utils.cpp_extension and runs the UNMODIFIED reference
// pychain/loss.py on top of it, so the C ABI is exercised from C++ exactly as the reference would use it.
//
// There is deliberately no CPU branch: the engine is CUDA-only (the reference keeps its own CPU path).
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <map>
#include <mutex>
#include <vector>

#include "chain_b200.h"

namespace {

void check_status(int st, const char *what) {
  TORCH_CHECK(st == 0, what, ": ", chain_b200_status_string(st), " (status ", st, ")");
}

bool is_shared(const torch::Tensor &t) { return t.dim() >= 1 && (t.size(0) == 1 || t.stride(0) == 0); }

// row 0 (shared) or the whole batch, contiguous, on `dev`
torch::Tensor graph_tensor(const torch::Tensor &t, bool shared, const torch::Device &dev, torch::ScalarType ty) {
  TORCH_CHECK(t.scalar_type() == ty, "graph tensor has the wrong dtype");
  TORCH_CHECK(shared || t.is_contiguous(), "graph tensors must be contiguous");  // CHECK_CONTIGUOUS, pychain.cc:24
  return (shared ? t.select(0, 0) : t).to(dev).contiguous();
}

// A denominator graph shared by the batch is packed once (chain_b200_den_prepare) and kept on the device,
// keyed by the storage of its forward transitions.
struct Packed {
  torch::Tensor blob_host, blob_dev;
  torch::Tensor keepalive;  // the source tensor: its data_ptr is the key
};
std::map<std::pair<const void *, int>, Packed> g_packed;
std::mutex g_mu;

}  // namespace

// pychain.cc:26-79
std::vector<torch::Tensor> ForwardBackward(
    torch::Tensor forward_transitions, torch::Tensor forward_transition_indices, torch::Tensor forward_transition_probs,
    torch::Tensor backward_transitions, torch::Tensor backward_transition_indices,
    torch::Tensor backward_transition_probs, torch::Tensor leaky_probs, torch::Tensor initial_probs,
    torch::Tensor final_probs, torch::Tensor start_state, torch::Tensor exp_nnet_output, torch::Tensor batch_sizes,
    torch::Tensor sequence_lengths, int num_states, float leaky_hmm_coefficient) {
  TORCH_CHECK(exp_nnet_output.is_cuda(), "chain_b200: exp_nnet_output must be a CUDA tensor (no CPU path)");
  TORCH_CHECK(exp_nnet_output.is_contiguous(), "exp_nnet_output must be contiguous");
  TORCH_CHECK(exp_nnet_output.scalar_type() == torch::kFloat, "exp_nnet_output must be float32");
  const c10::cuda::CUDAGuard guard(exp_nnet_output.device());
  const int B = exp_nnet_output.size(0), T = exp_nnet_output.size(1), D = exp_nnet_output.size(2);
  const int S = num_states, K = forward_transitions.size(1);
  const auto dev = exp_nnet_output.device();
  const bool shared = is_shared(forward_transitions);
  auto len = sequence_lengths.to(dev, torch::kLong).contiguous();
  auto objf = torch::empty({}, exp_nnet_output.options());
  auto grad = torch::empty_like(exp_nnet_output);
  auto ok = torch::empty({1}, exp_nnet_output.options().dtype(torch::kInt));
  void *stream = at::cuda::getCurrentCUDAStream().stream();

  if (shared && B > 1) {  // ChainGraphBatch(den_graph, B): pack once, run the shared-graph kernels
    auto host = [&](const torch::Tensor &t, torch::ScalarType ty) {
      TORCH_CHECK(t.scalar_type() == ty, "graph tensor has the wrong dtype");
      return t.select(0, 0).cpu().contiguous();
    };
    Packed *pk = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_mu);
      auto key = std::make_pair((const void *)forward_transitions.data_ptr(), (int)dev.index());
      auto it = g_packed.find(key);
      if (it == g_packed.end()) {
        auto ft = host(forward_transitions, torch::kInt), fi = host(forward_transition_indices, torch::kInt),
             fp = host(forward_transition_probs, torch::kFloat), bt = host(backward_transitions, torch::kInt),
             bi = host(backward_transition_indices, torch::kInt), bp = host(backward_transition_probs, torch::kFloat),
             lk = host(leaky_probs, torch::kFloat), in = host(initial_probs, torch::kFloat),
             fn = host(final_probs, torch::kFloat);
        const unsigned pflags = CHAIN_B200_PACK_G(chain_b200_den_group_width(S, D, B, dev.index()));
        const int64_t n = chain_b200_den_prepare_bytes(ft.data_ptr<int>(), fi.data_ptr<int>(), bt.data_ptr<int>(),
                                                       bi.data_ptr<int>(), S, K, D, pflags);
        TORCH_CHECK(n > 0, "chain_b200_den_prepare_bytes: ", chain_b200_status_string((int)n));
        Packed p;
        p.blob_host = torch::empty({n}, torch::dtype(torch::kByte));
        check_status(chain_b200_den_prepare(ft.data_ptr<int>(), fi.data_ptr<int>(), fp.data_ptr<float>(),
                                            bt.data_ptr<int>(), bi.data_ptr<int>(), bp.data_ptr<float>(),
                                            lk.data_ptr<float>(), in.data_ptr<float>(), fn.data_ptr<float>(), S, K, D,
                                            pflags, p.blob_host.data_ptr(), n),
                     "chain_b200_den_prepare");
        p.blob_dev = p.blob_host.to(dev);
        p.keepalive = forward_transitions;
        it = g_packed.emplace(key, std::move(p)).first;
      }
      pk = &it->second;
    }
    const size_t ws_bytes = chain_b200_den_workspace_bytes(pk->blob_host.data_ptr(), B, T);
    auto ws = torch::empty({(int64_t)ws_bytes}, exp_nnet_output.options().dtype(torch::kByte));
    check_status(chain_b200_den_forward_backward(pk->blob_host.data_ptr(), pk->blob_dev.data_ptr(), 1,
                                                 exp_nnet_output.data_ptr<float>(), len.data_ptr<int64_t>(), B, T, D,
                                                 leaky_hmm_coefficient, CHAIN_B200_X_IS_EXP, /*grad_scale*/ 1.0f, objf.data_ptr<float>(),
                                                 nullptr, grad.data_ptr<float>(), ok.data_ptr<int>(), nullptr,
                                                 ws.data_ptr(), ws_bytes, dev.index(), stream),
                 "chain_b200_den_forward_backward");
    return {objf, grad, ok.to(torch::kBool).cpu()};
  }

  const unsigned flags = CHAIN_B200_X_IS_EXP | (shared ? CHAIN_B200_GRAPH_SHARED : 0u);
  auto ft = graph_tensor(forward_transitions, shared, dev, torch::kInt),
       fi = graph_tensor(forward_transition_indices, shared, dev, torch::kInt),
       fp = graph_tensor(forward_transition_probs, shared, dev, torch::kFloat),
       bt = graph_tensor(backward_transitions, shared, dev, torch::kInt),
       bi = graph_tensor(backward_transition_indices, shared, dev, torch::kInt),
       bp = graph_tensor(backward_transition_probs, shared, dev, torch::kFloat),
       lk = graph_tensor(leaky_probs, shared, dev, torch::kFloat),
       in = graph_tensor(initial_probs, shared, dev, torch::kFloat),
       fn = graph_tensor(final_probs, shared, dev, torch::kFloat);
  const size_t ws_bytes = chain_b200_workspace_bytes(B, T, D, S, K, flags);
  auto ws = torch::empty({(int64_t)ws_bytes}, exp_nnet_output.options().dtype(torch::kByte));
  check_status(chain_b200_forward_backward(ft.data_ptr<int>(), fi.data_ptr<int>(), fp.data_ptr<float>(),
                                           bt.data_ptr<int>(), bi.data_ptr<int>(), bp.data_ptr<float>(),
                                           lk.data_ptr<float>(), in.data_ptr<float>(), fn.data_ptr<float>(),
                                           exp_nnet_output.data_ptr<float>(), len.data_ptr<int64_t>(), B, T, D, S, K,
                                           leaky_hmm_coefficient, flags, /*grad_scale*/ 1.0f, objf.data_ptr<float>(), nullptr,
                                           grad.data_ptr<float>(), ok.data_ptr<int>(), nullptr, ws.data_ptr(), ws_bytes,
                                           dev.index(), stream),
               "chain_b200_forward_backward");
  return {objf, grad, ok.to(torch::kBool).cpu()};
}

// pychain.cc:81-129
std::vector<torch::Tensor> ForwardBackwardLogDomain(
    torch::Tensor forward_transitions, torch::Tensor forward_transition_indices, torch::Tensor forward_transition_probs,
    torch::Tensor backward_transitions, torch::Tensor backward_transition_indices,
    torch::Tensor backward_transition_probs, torch::Tensor initial_probs, torch::Tensor final_probs,
    torch::Tensor start_state, torch::Tensor nnet_output, torch::Tensor batch_sizes, torch::Tensor sequence_lengths,
    int num_states) {
  TORCH_CHECK(nnet_output.is_cuda(), "chain_b200: nnet_output must be a CUDA tensor (no CPU path)");
  TORCH_CHECK(nnet_output.is_contiguous(), "nnet_output must be contiguous");
  TORCH_CHECK(nnet_output.scalar_type() == torch::kFloat, "nnet_output must be float32");
  const c10::cuda::CUDAGuard guard(nnet_output.device());
  const int B = nnet_output.size(0), T = nnet_output.size(1), D = nnet_output.size(2);
  const int S = num_states, K = forward_transitions.size(1);
  const auto dev = nnet_output.device();
  const bool shared = is_shared(forward_transitions);
  const unsigned flags = CHAIN_B200_GRAD_AS_LOG | (shared ? CHAIN_B200_GRAPH_SHARED : 0u);  // caller exp()s, loss.py:77
  auto ft = graph_tensor(forward_transitions, shared, dev, torch::kInt),
       fi = graph_tensor(forward_transition_indices, shared, dev, torch::kInt),
       fp = graph_tensor(forward_transition_probs, shared, dev, torch::kFloat),
       bt = graph_tensor(backward_transitions, shared, dev, torch::kInt),
       bi = graph_tensor(backward_transition_indices, shared, dev, torch::kInt),
       bp = graph_tensor(backward_transition_probs, shared, dev, torch::kFloat),
       in = graph_tensor(initial_probs, shared, dev, torch::kFloat),
       fn = graph_tensor(final_probs, shared, dev, torch::kFloat);
  auto len = sequence_lengths.to(dev, torch::kLong).contiguous();
  auto objf = torch::empty({}, nnet_output.options());
  auto grad = torch::empty_like(nnet_output);
  auto ok = torch::empty({1}, nnet_output.options().dtype(torch::kInt));
  const size_t ws_bytes = chain_b200_workspace_bytes(B, T, D, S, K, flags);
  auto ws = torch::empty({(int64_t)ws_bytes}, nnet_output.options().dtype(torch::kByte));
  check_status(chain_b200_forward_backward_log_domain(
                   ft.data_ptr<int>(), fi.data_ptr<int>(), fp.data_ptr<float>(), bt.data_ptr<int>(), bi.data_ptr<int>(),
                   bp.data_ptr<float>(), in.data_ptr<float>(), fn.data_ptr<float>(), nnet_output.data_ptr<float>(),
                   len.data_ptr<int64_t>(), B, T, D, S, K, flags, /*grad_scale*/ 1.0f, objf.data_ptr<float>(), nullptr,
                   grad.data_ptr<float>(), ok.data_ptr<int>(), nullptr, ws.data_ptr(), ws_bytes, dev.index(),
                   at::cuda::getCurrentCUDAStream().stream()),
               "chain_b200_forward_backward_log_domain");
  return {objf, grad, ok.to(torch::kBool).cpu()};
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("forward_backward", &ForwardBackward, "LF-MMI forward-backward (chain_b200 C ABI)");
  m.def("forward_backward_log_domain", &ForwardBackwardLogDomain, "log-domain forward-backward (chain_b200 C ABI)");
  m.def("set_verbose_level", [](int level) { chain_b200_set_verbose_level(level); });
}
