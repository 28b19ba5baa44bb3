// torch_binding.cpp -- the Python module `torch_asg_native`: the seven functions the reference registers
// in native/extension.cpp:15-29, same names, same positional arguments, same tuple returns, implemented as
// thin calls into the C ABI of libasg_b200.so (include/asg_b200.h).  torch is plumbing here: it owns the
// device memory and the current stream, nothing else.
//
// What the returned "alpha / beta / gamma / path_contrib" tensors hold differs from the reference by design
// (asg.py only looks at alpha.shape / full_gamma.shape -- asg.py:28,50,88 -- and hands everything else
// straight back to our own backward):
//   alpha / full_gamma         scaled real-domain FCC alpha: a [T,N,L] strided view into the frame-record array
//   aligned alpha / gamma      block-floating-point FAC alpha: a [T,N,S] strided view into the same array
//   beta                       the matching beta part of the records
//   path_contrib               opaque uint8 blobs (the context blob / the frame-record array, asg_layout_t) -- NOT [T-1,N,L,L]
// There is no CPU path: CPU tensors raise.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "../../include/asg_b200.h"

namespace {

using at::Tensor;

int dtype_code(const Tensor &t, const char *what) {
    if (t.scalar_type() == at::kFloat) return ASG_F32;
    if (t.scalar_type() == at::kDouble) return ASG_F64;
    TORCH_CHECK(false, "expected scalar type Float or Double but found ", t.scalar_type(), " for ", what);
    return -1;
}

void check_cuda(const Tensor &t, const char *what) {
    TORCH_CHECK(t.is_cuda(), "torch_asg_native (B200 build): '", what,
                "' must be a CUDA tensor -- this extension has no CPU implementation and no fallback");
}

// lengths / targets: int64 only, like the reference (utils.cpp:28,46; force_aligned_lattice.cpp:31-33).
// numel >= 0: the tensor must hold exactly that many elements (a short lengths / targets tensor would be read out of bounds)
Tensor as_device_long(const Tensor &t, const Tensor &like, const char *what, int64_t numel = -1) {
    TORCH_CHECK(t.scalar_type() == at::kLong, "expected scalar type Long but found ", t.scalar_type(), " for ", what);
    TORCH_CHECK(numel < 0 || t.numel() == numel, what, " must hold ", numel, " elements, got ", t.numel());
    Tensor r = t;
    if (r.device() != like.device()) r = r.to(like.device(), /*non_blocking=*/true);
    return r.contiguous();
}

void check_rc(int rc) { TORCH_CHECK(rc == 0, "asg_b200: ", asg_last_error(), " (code ", rc, ")"); }

asg_layout_t layout_of(int dtype, int64_t T, int64_t N, int64_t L, int64_t S) {
    asg_layout_t lay;
    check_rc(asg_layout(dtype, T, N, L, S, &lay));
    return lay;
}

Tensor bytes_like(const Tensor &like, size_t nbytes) {
    return at::empty({(int64_t)nbytes}, like.options().dtype(at::kByte).requires_grad(false));
}

// [T,N,W] view (what asg.py inspects the shape of) of one part of the frame records held by the byte tensor `state`
Tensor rec_view(const Tensor &state, const asg_layout_t &lay, size_t part, at::ScalarType dt, int64_t T, int64_t N, int64_t W) {
    const int64_t eb = (int64_t)lay.elem_bytes, rs = (int64_t)lay.rec_bytes / eb;
    return state.view(dt).as_strided({T, N, W}, {rs, lay.tp * rs, 1}, (int64_t)part / eb);
}
// the base of the frame-record array a view produced by rec_view() points into (or raises)
void *rec_base(const Tensor &v, const asg_layout_t &lay, size_t part, const char *what) {
    const int64_t eb = (int64_t)lay.elem_bytes, rs = (int64_t)lay.rec_bytes / eb;
    TORCH_CHECK(v.dim() == 3 && v.stride(2) == 1 && v.stride(0) == rs && v.stride(1) == lay.tp * rs &&
                    v.storage_offset() == (int64_t)part / eb && (size_t)v.storage().nbytes() >= lay.state_bytes,
                what, " is not a state tensor produced by this extension's forward");
    return static_cast<char *>(v.data_ptr()) - part;
}

void check_inputs(const Tensor &inputs, const Tensor &transition, int64_t T, int64_t N, int64_t L) {
    check_cuda(inputs, "inputs");
    check_cuda(transition, "transition");
    TORCH_CHECK(inputs.dim() == 3 && inputs.size(0) == T && inputs.size(1) == N && inputs.size(2) == L,
                "inputs must be [T,N,L] = [", T, ",", N, ",", L, "]");
    TORCH_CHECK(transition.dim() == 2 && transition.size(0) == L && transition.size(1) == L, "transition must be [L,L]");
    TORCH_CHECK(transition.scalar_type() == inputs.scalar_type(), "inputs and transition must share a dtype");
    TORCH_CHECK(transition.device() == inputs.device(), "inputs and transition must share a device");
}

cudaStream_t current_stream() { return c10::cuda::getCurrentCUDAStream().stream(); }

// ------------------------------------------------------------------------------------------------
// reference: fully_connected_forward, fully_connected_lattice.h:41-49 / .cpp:66-91
std::tuple<Tensor, Tensor, Tensor, Tensor> fully_connected_forward(Tensor &inputs, Tensor &transition,
                                                                   Tensor &input_lengths, int64_t T, int64_t N, int64_t L) {
    check_inputs(inputs, transition, T, N, L);
    c10::cuda::CUDAGuard guard(inputs.device());
    const int dt = dtype_code(inputs, "inputs");
    Tensor il = as_device_long(input_lengths, inputs, "input_lengths", N);
    const asg_layout_t lay = layout_of(dt, T, N, L, 1);
    auto opt = inputs.options().requires_grad(false);
    Tensor scores = at::empty({N}, opt);
    Tensor state = bytes_like(inputs, lay.state_bytes);
    Tensor ctx = bytes_like(inputs, lay.ctx_bytes);
    check_rc(asg_scan(dt, ASG_ROLE_FCC_ALPHA | ASG_ROLE_FCC_BETA, inputs.data_ptr(), inputs.stride(0), inputs.stride(1),
                      inputs.stride(2), transition.data_ptr(), transition.stride(0), transition.stride(1), nullptr, 0, 0,
                      il.data_ptr<int64_t>(), nullptr, T, N, L, 1, state.data_ptr(), ctx.data_ptr(), scores.data_ptr(), nullptr,
                      current_stream()));
    const auto st = inputs.scalar_type();
    return {scores, rec_view(state, lay, lay.rec_fcc_alpha, st, T, N, L), rec_view(state, lay, lay.rec_fcc_beta, st, T, N, L), ctx};
}

// reference: fully_connected_backward, fully_connected_lattice.h:51-60 / .cpp:94-104
std::tuple<Tensor, Tensor> fully_connected_backward(Tensor &grad_out, Tensor &alpha, Tensor &beta, Tensor &path_contrib,
                                                    int64_t T, int64_t N, int64_t L) {
    check_cuda(alpha, "alpha");
    c10::cuda::CUDAGuard guard(alpha.device());
    const int dt = dtype_code(alpha, "alpha");
    const asg_layout_t lay = layout_of(dt, T, N, L, 1);
    TORCH_CHECK((size_t)path_contrib.numel() == lay.ctx_bytes && path_contrib.scalar_type() == at::kByte,
                "path_contrib is not a context blob produced by this extension's fully_connected_forward");
    void *state = rec_base(alpha, lay, lay.rec_fcc_alpha, "alpha");
    Tensor g = grad_out.to(alpha.options());
    auto opt = alpha.options().requires_grad(false);
    Tensor grad_x = at::empty({T, N, L}, opt), grad_A = at::empty({L, L}, opt);
    Tensor scratch = bytes_like(alpha, lay.grad_scratch_bytes);
    check_rc(asg_grad(dt, ASG_PART_FCC, g.data_ptr(), g.dim() ? g.stride(0) : 0, nullptr, 0, state, path_contrib.data_ptr(),
                      nullptr, 0, 0, T, N, L, 1, grad_x.data_ptr(), grad_A.data_ptr(), scratch.data_ptr(), current_stream()));
    return {grad_A, grad_x};
}

// reference: force_aligned_forward, force_aligned_lattice.h:42-53 / .cpp:266-319
std::tuple<Tensor, Tensor, Tensor, Tensor> force_aligned_forward(Tensor &inputs, Tensor &targets, Tensor &transition,
                                                                 Tensor &input_lengths, Tensor &target_lengths, int64_t T,
                                                                 int64_t N, int64_t L, int64_t S) {
    check_inputs(inputs, transition, T, N, L);
    c10::cuda::CUDAGuard guard(inputs.device());
    const int dt = dtype_code(inputs, "inputs");
    Tensor y = as_device_long(targets, inputs, "targets", N * S);
    Tensor il = as_device_long(input_lengths, inputs, "input_lengths", N);
    Tensor tl = as_device_long(target_lengths, inputs, "target_lengths", N);
    TORCH_CHECK(y.dim() == 2 && y.size(0) == N && y.size(1) == S, "targets must be [N,S]");
    const asg_layout_t lay = layout_of(dt, T, N, L, S);
    auto opt = inputs.options().requires_grad(false);
    Tensor scores = at::empty({N}, opt);
    Tensor state = bytes_like(inputs, lay.state_bytes);
    Tensor ctx = bytes_like(inputs, lay.ctx_bytes);
    check_rc(asg_scan(dt, ASG_ROLE_FAC_ALPHA | ASG_ROLE_FAC_BETA, inputs.data_ptr(), inputs.stride(0), inputs.stride(1),
                      inputs.stride(2), transition.data_ptr(), transition.stride(0), transition.stride(1),
                      y.data_ptr<int64_t>(), y.stride(0), y.stride(1), il.data_ptr<int64_t>(), tl.data_ptr<int64_t>(), T, N, L,
                      S, state.data_ptr(), ctx.data_ptr(), nullptr, scores.data_ptr(), current_stream()));
    const auto st = inputs.scalar_type();
    return {scores, rec_view(state, lay, lay.rec_fac_alpha, st, T, N, S), rec_view(state, lay, lay.rec_fac_beta, st, T, N, S), ctx};
}

// reference: force_aligned_backward, force_aligned_lattice.h:56-69 / .cpp:321-356
std::tuple<Tensor, Tensor> force_aligned_backward(Tensor &grad_out, Tensor &alpha, Tensor &beta, Tensor &path_contrib,
                                                  Tensor &targets, Tensor &input_lengths, Tensor &target_lengths, int64_t T,
                                                  int64_t N, int64_t L, int64_t S) {
    check_cuda(alpha, "alpha");
    c10::cuda::CUDAGuard guard(alpha.device());
    const int dt = dtype_code(alpha, "alpha");
    const asg_layout_t lay = layout_of(dt, T, N, L, S);
    TORCH_CHECK((size_t)path_contrib.numel() == lay.ctx_bytes && path_contrib.scalar_type() == at::kByte,
                "path_contrib is not a context blob produced by this extension's force_aligned_forward");
    void *state = rec_base(alpha, lay, lay.rec_fac_alpha, "alpha");
    Tensor y = as_device_long(targets, alpha, "targets", N * S);
    Tensor g = grad_out.to(alpha.options());
    auto opt = alpha.options().requires_grad(false);
    Tensor grad_x = at::empty({T, N, L}, opt), grad_A = at::empty({L, L}, opt);
    Tensor scratch = bytes_like(alpha, lay.grad_scratch_bytes);
    check_rc(asg_grad(dt, ASG_PART_FAC, nullptr, 0, g.data_ptr(), g.dim() ? g.stride(0) : 0, state, path_contrib.data_ptr(),
                      y.data_ptr<int64_t>(), y.stride(0), y.stride(1), T, N, L, S, grad_x.data_ptr(), grad_A.data_ptr(),
                      scratch.data_ptr(), current_stream()));
    return {grad_A, grad_x};
}

// reference: fast_asg_gpu_forward_only, streamlined_fast_gpu.h:17-28 / .cpp:24-94
Tensor fast_asg_gpu_forward_only(Tensor &inputs, Tensor &targets, Tensor &transition, Tensor &input_lengths,
                                 Tensor &target_lengths, int64_t T, int64_t N, int64_t L, int64_t S) {
    check_inputs(inputs, transition, T, N, L);
    c10::cuda::CUDAGuard guard(inputs.device());
    const int dt = dtype_code(inputs, "inputs");
    Tensor y = as_device_long(targets, inputs, "targets", N * S);
    Tensor il = as_device_long(input_lengths, inputs, "input_lengths", N);
    Tensor tl = as_device_long(target_lengths, inputs, "target_lengths", N);
    TORCH_CHECK(y.dim() == 2 && y.size(0) == N && y.size(1) == S, "targets must be [N,S]");
    const asg_layout_t lay = layout_of(dt, T, N, L, S);
    auto opt = inputs.options().requires_grad(false);
    Tensor full = at::empty({N}, opt), aligned = at::empty({N}, opt);
    // forward-only keeps no state: no frame records, only the small context blob (E, lengths, scalars)
    Tensor ctx = bytes_like(inputs, lay.ctx_bytes);
    check_rc(asg_scan(dt, ASG_ROLE_FORWARD_ONLY, inputs.data_ptr(), inputs.stride(0), inputs.stride(1), inputs.stride(2),
                      transition.data_ptr(), transition.stride(0), transition.stride(1), y.data_ptr<int64_t>(), y.stride(0),
                      y.stride(1), il.data_ptr<int64_t>(), tl.data_ptr<int64_t>(), T, N, L, S, nullptr, ctx.data_ptr(),
                      full.data_ptr(), aligned.data_ptr(), current_stream()));
    return full - aligned;
}

// reference: fast_asg_gpu_forward, streamlined_fast_gpu.h:30-48 / .cpp:104-230
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> fast_asg_gpu_forward(Tensor &inputs, Tensor &targets,
                                                                                Tensor &transition, Tensor &input_lengths,
                                                                                Tensor &target_lengths, int64_t T, int64_t N,
                                                                                int64_t L, int64_t S) {
    check_inputs(inputs, transition, T, N, L);
    c10::cuda::CUDAGuard guard(inputs.device());
    const int dt = dtype_code(inputs, "inputs");
    Tensor y = as_device_long(targets, inputs, "targets", N * S);
    Tensor il = as_device_long(input_lengths, inputs, "input_lengths", N);
    Tensor tl = as_device_long(target_lengths, inputs, "target_lengths", N);
    TORCH_CHECK(y.dim() == 2 && y.size(0) == N && y.size(1) == S, "targets must be [N,S]");
    const asg_layout_t lay = layout_of(dt, T, N, L, S);
    auto opt = inputs.options().requires_grad(false);
    Tensor full = at::empty({N}, opt), aligned = at::empty({N}, opt);
    // the frame records travel in the "aligned_path_contrib" slot, the context blob in "full_path_contrib"
    Tensor state = bytes_like(inputs, lay.state_bytes);
    Tensor ctx = bytes_like(inputs, lay.ctx_bytes);
    check_rc(asg_scan(dt, ASG_ROLE_ALL, inputs.data_ptr(), inputs.stride(0), inputs.stride(1), inputs.stride(2),
                      transition.data_ptr(), transition.stride(0), transition.stride(1), y.data_ptr<int64_t>(), y.stride(0),
                      y.stride(1), il.data_ptr<int64_t>(), tl.data_ptr<int64_t>(), T, N, L, S, state.data_ptr(), ctx.data_ptr(),
                      full.data_ptr(), aligned.data_ptr(), current_stream()));
    const auto st = inputs.scalar_type();
    return {full, aligned, rec_view(state, lay, lay.rec_fcc_alpha, st, T, N, L), rec_view(state, lay, lay.rec_fac_alpha, st, T, N, S), ctx, state};
}

// reference: fast_asg_gpu_backward, streamlined_fast_gpu.h:50-68 / .cpp:236-297
std::tuple<Tensor, Tensor> fast_asg_gpu_backward(Tensor &grad_full, Tensor &grad_aligned, Tensor &full_gamma,
                                                 Tensor &aligned_gamma, Tensor &full_path_contrib,
                                                 Tensor &aligned_path_contrib, Tensor &targets, Tensor &input_lengths,
                                                 Tensor &target_lengths, int64_t T, int64_t N, int64_t L, int64_t S) {
    check_cuda(full_gamma, "full_gamma");
    c10::cuda::CUDAGuard guard(full_gamma.device());
    const int dt = dtype_code(full_gamma, "full_gamma");
    const asg_layout_t lay = layout_of(dt, T, N, L, S);
    TORCH_CHECK((size_t)full_path_contrib.numel() == lay.ctx_bytes && full_path_contrib.scalar_type() == at::kByte &&
                    (size_t)aligned_path_contrib.numel() == lay.state_bytes && aligned_path_contrib.scalar_type() == at::kByte,
                "saved tensors were not produced by this extension's fast_asg_gpu_forward");
    TORCH_CHECK(rec_base(full_gamma, lay, lay.rec_fcc_alpha, "full_gamma") == aligned_path_contrib.data_ptr() &&
                    rec_base(aligned_gamma, lay, lay.rec_fac_alpha, "aligned_gamma") == aligned_path_contrib.data_ptr(),
                "saved tensors were not produced by the same fast_asg_gpu_forward call");
    Tensor y = as_device_long(targets, full_gamma, "targets", N * S);
    Tensor gf = grad_full.to(full_gamma.options()), ga = grad_aligned.to(full_gamma.options());
    auto opt = full_gamma.options().requires_grad(false);
    Tensor grad_x = at::empty({T, N, L}, opt), grad_A = at::empty({L, L}, opt);
    Tensor scratch = bytes_like(full_gamma, lay.grad_scratch_bytes);
    check_rc(asg_grad(dt, ASG_PART_BOTH, gf.data_ptr(), gf.dim() ? gf.stride(0) : 0, ga.data_ptr(), ga.dim() ? ga.stride(0) : 0,
                      aligned_path_contrib.data_ptr(), full_path_contrib.data_ptr(), y.data_ptr<int64_t>(), y.stride(0), y.stride(1),
                      T, N, L, S, grad_x.data_ptr(), grad_A.data_ptr(), scratch.data_ptr(), current_stream()));
    return {grad_A, grad_x};
}

// not in the reference: the range report of a context blob (asg_range_flags): int32 [N]
Tensor range_flags(Tensor &state, Tensor &ctx, bool is_double, int64_t T, int64_t N, int64_t L, int64_t S) {
    check_cuda(ctx, "ctx");
    check_cuda(state, "state");
    const asg_layout_t lay = layout_of(is_double ? ASG_F64 : ASG_F32, T, N, L, S);
    TORCH_CHECK((size_t)ctx.numel() == lay.ctx_bytes && ctx.scalar_type() == at::kByte, "not a context blob for these dimensions");
    // forward_only results carry no states (an empty tensor): only the flags the scans raise themselves are reported then
    const bool has_state = (size_t)state.numel() == lay.state_bytes && state.scalar_type() == at::kByte;
    TORCH_CHECK(has_state || state.numel() == 0, "not a state array for these dimensions");
    c10::cuda::CUDAGuard guard(ctx.device());
    Tensor out = at::empty({N}, ctx.options().dtype(at::kInt));
    check_rc(asg_range_flags(is_double ? ASG_F64 : ASG_F32, has_state ? state.data_ptr() : nullptr, ctx.data_ptr(), T, N, L, S, out.data_ptr<int32_t>(), current_stream()));
    return out;
}

// ------------------------------------------------------------------------------------------------
// Workspace: every buffer one (dtype, T, N, L, S) problem needs, allocated ONCE, so a step is pure kernel launches
// (asg_forward / asg_backward / asg_step allocate nothing and never synchronise) and can be captured in a CUDA graph.
// torch owns the memory (the fused buffer is what torch.distributed all-reduces); the arithmetic is the C ABI's.
struct Workspace {
    int dt;
    int64_t T, N, L, S;
    asg_layout_t lay;
    Tensor state, ctx, scratch, full, aligned, losses, loss_tail, grad_x, fused, gout;

    Workspace(bool is_double, int64_t device, int64_t T_, int64_t N_, int64_t L_, int64_t S_) : T(T_), N(N_), L(L_), S(S_) {
        dt = is_double ? ASG_F64 : ASG_F32;
        lay = layout_of(dt, T, N, L, S);
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
        auto opt = at::TensorOptions().device(at::kCUDA, (c10::DeviceIndex)device).dtype(is_double ? at::kDouble : at::kFloat);
        auto bopt = opt.dtype(at::kByte);
        state = at::empty({(int64_t)lay.state_bytes}, bopt);
        ctx = at::zeros({(int64_t)lay.ctx_bytes}, bopt);
        scratch = at::empty({(int64_t)lay.grad_scratch_bytes}, bopt);
        full = at::zeros({N}, opt);
        aligned = at::zeros({N}, opt);
        losses = at::zeros({N}, opt);
        loss_tail = at::zeros({3}, opt);
        grad_x = at::zeros({T, N, L}, opt);
        fused = at::zeros({L * L + 2}, opt);
        gout = at::ones({1}, opt);
    }

    void check(const Tensor &x, const Tensor &A, const Tensor &y, const Tensor &il, const Tensor &tl) const {
        check_inputs(x, A, T, N, L);
        TORCH_CHECK(dtype_code(x, "inputs") == dt, "inputs dtype differs from the workspace's");
        TORCH_CHECK(x.device() == state.device(), "inputs are not on the workspace's device");
        for (const Tensor *t : {&y, &il, &tl}) {
            TORCH_CHECK(t->scalar_type() == at::kLong, "expected scalar type Long but found ", t->scalar_type());
            TORCH_CHECK(t->device() == x.device(), "targets and lengths must be on the inputs' device (the workspace path never copies)");
        }
        TORCH_CHECK(y.dim() == 2 && y.size(0) == N && y.size(1) == S, "targets must be [N,S]");
        TORCH_CHECK(il.dim() == 1 && il.size(0) == N && il.is_contiguous() && tl.dim() == 1 && tl.size(0) == N && tl.is_contiguous(),
                    "lengths must be contiguous [N]");
    }

    void forward(Tensor &x, Tensor &A, Tensor &y, Tensor &il, Tensor &tl, bool forward_only, double scale) {
        check(x, A, y, il, tl);
        c10::cuda::CUDAGuard guard(x.device());
        check_rc(asg_forward(dt, forward_only ? 1 : 0, x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), A.data_ptr(), A.stride(0),
                             A.stride(1), y.data_ptr<int64_t>(), y.stride(0), y.stride(1), il.data_ptr<int64_t>(), tl.data_ptr<int64_t>(),
                             T, N, L, S, scale, state.data_ptr(), ctx.data_ptr(), full.data_ptr(), aligned.data_ptr(), losses.data_ptr(),
                             loss_tail.data_ptr(), current_stream()));
    }
    // upstream gradient of utterance n's loss = scale * g[n] (g: [N]), or scale * g[0] for a one-element g
    void backward(Tensor &g, double scale, Tensor &y) {
        TORCH_CHECK(g.is_cuda() && g.device() == state.device() && dtype_code(g, "grad") == dt && (g.numel() == 1 || g.numel() == N),
                    "grad must be a device tensor of 1 or N elements in the workspace's dtype");
        c10::cuda::CUDAGuard guard(state.device());
        check_rc(asg_backward(dt, g.data_ptr(), g.numel() == 1 ? 0 : (g.dim() ? g.stride(g.dim() - 1) : 0), scale, state.data_ptr(),
                              ctx.data_ptr(), y.data_ptr<int64_t>(), y.stride(0), y.stride(1), T, N, L, S, grad_x.data_ptr(),
                              fused.data_ptr(), scratch.data_ptr(), current_stream()));
    }
    Tensor flags() { return range_flags(state, ctx, dt == ASG_F64, T, N, L, S); }
    void step(Tensor &x, Tensor &A, Tensor &y, Tensor &il, Tensor &tl, double scale) {
        check(x, A, y, il, tl);
        c10::cuda::CUDAGuard guard(x.device());
        check_rc(asg_step(dt, x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), A.data_ptr(), A.stride(0), A.stride(1),
                          y.data_ptr<int64_t>(), y.stride(0), y.stride(1), il.data_ptr<int64_t>(), tl.data_ptr<int64_t>(), T, N, L, S,
                          scale, state.data_ptr(), ctx.data_ptr(), scratch.data_ptr(), full.data_ptr(), aligned.data_ptr(),
                          losses.data_ptr(), grad_x.data_ptr(), fused.data_ptr(), current_stream()));
    }
};

// not in the reference: the exchange step of a batch-sharded job over NVLink peer memory (asg_peer_*)
struct PeerComm {
    asg_peer_t *p = nullptr;
    int64_t device;
    std::string handle_;
    PeerComm(int64_t rank, int64_t world, int64_t max_bytes, int64_t device_) : device(device_) {
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
        char h[ASG_PEER_HANDLE_BYTES];
        check_rc(asg_peer_create((int)rank, (int)world, (size_t)max_bytes, &p, h));
        handle_.assign(h, sizeof(h));
    }
    ~PeerComm() { if (p) { c10::cuda::CUDAGuard guard((c10::DeviceIndex)device); asg_peer_destroy(p); } }
    PeerComm(const PeerComm &) = delete;
    pybind11::bytes handle() const { return pybind11::bytes(handle_); }
    void connect(const std::string &handles) {
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
        check_rc(asg_peer_connect(p, handles.data()));
    }
    void allreduce(Tensor &buf) {
        check_cuda(buf, "buf");
        TORCH_CHECK(buf.is_contiguous() && (buf.scalar_type() == at::kFloat || buf.scalar_type() == at::kDouble) && buf.get_device() == device,
                    "peer all-reduce takes a contiguous float32/float64 tensor on the communicator's device");
        c10::cuda::CUDAGuard guard(buf.device());
        check_rc(asg_peer_allreduce(p, buf.scalar_type() == at::kDouble ? ASG_F64 : ASG_F32, buf.data_ptr(), buf.numel(), current_stream()));
    }
    std::pair<uint64_t, uint64_t> status() {
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
        uint64_t e = 0, t = 0;
        check_rc(asg_peer_status(p, &e, &t));
        return {e, t};
    }
};

// not in the reference: the backward correction of the fused log_softmax + ASG path (asg_logits_backward_fix)
void logits_backward_fix(Tensor &z, Tensor &aligned, Tensor &g, double scale, Tensor &input_lengths, Tensor &grad_x) {
    check_cuda(z, "inputs");
    TORCH_CHECK(z.dim() == 3 && grad_x.dim() == 3 && grad_x.is_contiguous() && grad_x.sizes() == z.sizes() && grad_x.scalar_type() == z.scalar_type(),
                "grad_x must be a contiguous tensor shaped like inputs");
    const int64_t T = z.size(0), N = z.size(1), L = z.size(2);
    TORCH_CHECK(aligned.numel() == N && aligned.is_contiguous() && aligned.scalar_type() == z.scalar_type() && aligned.device() == z.device(),
                "aligned scores must be [N] in the inputs' dtype");
    TORCH_CHECK(g.numel() == 1 || g.numel() == N, "grad must hold 1 or N elements");
    c10::cuda::CUDAGuard guard(z.device());
    Tensor il = as_device_long(input_lengths, z, "input_lengths", N);
    Tensor gg = g.to(z.options()).contiguous();
    check_rc(asg_logits_backward_fix(dtype_code(z, "inputs"), z.data_ptr(), z.stride(0), z.stride(1), z.stride(2), aligned.data_ptr(), gg.data_ptr(),
                                     gg.numel() == 1 ? 0 : 1, scale, il.data_ptr<int64_t>(), T, N, L, grad_x.data_ptr(), current_stream()));
}

// not in the reference (README.md:33 lists it as to do): Viterbi forced alignment, asg_viterbi
std::tuple<Tensor, Tensor> viterbi_align(Tensor &inputs, Tensor &targets, Tensor &transition, Tensor &input_lengths, Tensor &target_lengths) {
    TORCH_CHECK(inputs.dim() == 3 && targets.dim() == 2 && targets.size(0) == inputs.size(1), "inputs [T,N,L], targets [N,S]");
    const int64_t T = inputs.size(0), N = inputs.size(1), L = inputs.size(2), S = targets.size(1);
    check_inputs(inputs, transition, T, N, L);
    c10::cuda::CUDAGuard guard(inputs.device());
    const int dt = dtype_code(inputs, "inputs");
    Tensor y = as_device_long(targets, inputs, "targets", N * S);
    Tensor il = as_device_long(input_lengths, inputs, "input_lengths", N);
    Tensor tl = as_device_long(target_lengths, inputs, "target_lengths", N);
    Tensor scores = at::empty({N}, inputs.options().requires_grad(false));
    Tensor path = at::empty({N, T}, inputs.options().dtype(at::kLong).requires_grad(false));
    Tensor scratch = bytes_like(inputs, asg_viterbi_scratch_bytes(T, N, S));
    check_rc(asg_viterbi(dt, inputs.data_ptr(), inputs.stride(0), inputs.stride(1), inputs.stride(2), transition.data_ptr(), transition.stride(0),
                         transition.stride(1), y.data_ptr<int64_t>(), y.stride(0), y.stride(1), il.data_ptr<int64_t>(), tl.data_ptr<int64_t>(),
                         T, N, L, S, scores.data_ptr(), path.data_ptr<int64_t>(), scratch.data_ptr(), current_stream()));
    return {scores, path};
}

// not in the reference: replabel packing / unpacking of integer targets on the device (asg_pack_replabels / asg_unpack_replabels)
std::tuple<Tensor, Tensor> replabels(Tensor &targets, Tensor &target_lengths, int64_t first_rep_label, int64_t max_reps, int64_t fill,
                                     bool pack, int64_t S_out) {
    check_cuda(targets, "targets");
    TORCH_CHECK(targets.dim() == 2 && targets.scalar_type() == at::kLong, "targets must be an int64 [N,S] tensor");
    const int64_t N = targets.size(0), S = targets.size(1);
    c10::cuda::CUDAGuard guard(targets.device());
    Tensor tl = as_device_long(target_lengths, targets, "target_lengths", N);
    if (pack) S_out = S;
    Tensor out = at::empty({N, S_out}, targets.options());
    Tensor len = at::empty({N}, targets.options());
    if (pack)
        check_rc(asg_pack_replabels(targets.data_ptr<int64_t>(), targets.stride(0), targets.stride(1), tl.data_ptr<int64_t>(), N, S, first_rep_label,
                                    max_reps, fill, out.data_ptr<int64_t>(), len.data_ptr<int64_t>(), current_stream()));
    else
        check_rc(asg_unpack_replabels(targets.data_ptr<int64_t>(), targets.stride(0), targets.stride(1), tl.data_ptr<int64_t>(), N, S, first_rep_label,
                                      max_reps, fill, out.data_ptr<int64_t>(), S_out, len.data_ptr<int64_t>(), current_stream()));
    return {out, len};
}

// ---- extras (not in the reference): introspection for tests and bench.py ------------------------
Tensor ctx_scalars(Tensor &ctx, bool is_double, int64_t T, int64_t N, int64_t L, int64_t S) {
    const asg_layout_t lay = layout_of(is_double ? ASG_F64 : ASG_F32, T, N, L, S);
    TORCH_CHECK((size_t)ctx.numel() == lay.ctx_bytes, "not a context blob for these dimensions");
    Tensor view = at::from_blob(static_cast<char *>(ctx.data_ptr()) + lay.off_scalars, {N, 8},
                                ctx.options().dtype(at::kDouble));
    return view.clone();
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "torch_asg_native -- B200 (sm_100a) implementation of zh217/torch-asg's native boundary";
    m.def("fully_connected_forward", &fully_connected_forward, "fully-connected forward");
    m.def("fully_connected_backward", &fully_connected_backward, "fully-connected backward");
    m.def("force_aligned_forward", &force_aligned_forward, "force-aligned forward");
    m.def("force_aligned_backward", &force_aligned_backward, "force-aligned backward");
    m.def("fast_asg_gpu_forward_only", &fast_asg_gpu_forward_only, "fast ASG GPU forward only");
    m.def("fast_asg_gpu_forward", &fast_asg_gpu_forward, "fast ASG GPU forward");
    m.def("fast_asg_gpu_backward", &fast_asg_gpu_backward, "fast ASG GPU backward");
    pybind11::class_<Workspace>(m, "Workspace",
                                "persistent buffers of one (dtype, T, N, L, S) problem; forward/backward/step allocate nothing "
                                "(CUDA-graph capturable).  Not in the reference: the host-independent step of this build.")
        .def(pybind11::init<bool, int64_t, int64_t, int64_t, int64_t, int64_t>(), pybind11::arg("is_double"), pybind11::arg("device"),
             pybind11::arg("T"), pybind11::arg("N"), pybind11::arg("L"), pybind11::arg("S"))
        .def("forward", &Workspace::forward)
        .def("backward", &Workspace::backward)
        .def("step", &Workspace::step)
        .def("range_flags", &Workspace::flags)
        .def_readonly("full_scores", &Workspace::full)
        .def_readonly("aligned_scores", &Workspace::aligned)
        .def_readonly("losses", &Workspace::losses)
        .def_readonly("loss_tail", &Workspace::loss_tail)
        .def_readonly("grad_x", &Workspace::grad_x)
        .def_readonly("fused", &Workspace::fused)
        .def_readonly("gout", &Workspace::gout)
        .def_readonly("state", &Workspace::state)
        .def_readonly("ctx", &Workspace::ctx)
        .def_readonly("T", &Workspace::T).def_readonly("N", &Workspace::N).def_readonly("L", &Workspace::L).def_readonly("S", &Workspace::S);
    pybind11::class_<PeerComm>(m, "PeerComm", "exchange buffers of a batch-sharded job over NVLink peer memory (asg_peer_*)")
        .def(pybind11::init<int64_t, int64_t, int64_t, int64_t>(), pybind11::arg("rank"), pybind11::arg("world"), pybind11::arg("max_bytes"),
             pybind11::arg("device"))
        .def("handle", &PeerComm::handle)
        .def("connect", &PeerComm::connect)
        .def("allreduce", &PeerComm::allreduce)
        .def("status", &PeerComm::status);
    m.def("range_flags", &range_flags, "per-utterance range report of (state, ctx) (asg_range_flags), int32 [N] (not in the reference)");
    m.def("replabels", &replabels, "replabel packing (pack=True) / unpacking of int64 targets: (out[N,S_out], lengths[N]) (not in the reference)");
    m.def("viterbi_align", &viterbi_align, "Viterbi forced alignment: (scores[N], path[N,T]) (not in the reference)");
    m.def("logits_backward_fix", &logits_backward_fix,
          "fused log_softmax + ASG: in-place correction of grad_x for infeasible utterances (not in the reference)");
    m.def("_ctx_scalars", &ctx_scalars, "per-utterance float64 bookkeeping [N,8] of a context blob");
    m.def("_launch_count", []() { return asg_launch_count(); }, "kernels launched by libasg_b200 so far");
    m.def("_version", []() { return std::string(asg_version()); });
    m.attr("IS_B200_BUILD") = true;
}
