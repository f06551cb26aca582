// Executor side of the tensor-core fused blocks (tc_block.cuh), f32 build only.
//
// With option "tc" the scheduler grows windows of at most 6 positions (scheduler.hpp, tile_bits = 6, low_bits = 0).
// A window whose gates are all unitary kinds runs as ONE dense 64 x 64 block:
//
//   forward : W = U_m ... U_1 (host, double) -> bf16 slice image -> k_tc_block_fwd            (2*S of HBM traffic)
//   reverse : ONE fused sweep (tc_rev.cuh: k_tc_block_rev, 4*S): state <- W^dagger state, adjoint <- W^T adjoint,
//             H += adjoint (x) state, and G_W = H conj(W) on the host.  (Option tc_rev = 0 keeps the three-sweep form:
//             k_tc_block_fwd, k_tc_block_grad, k_tc_block_fwd, 6*S.)
//             Once per backward() call, the chain rule from the block gradients G_W to the member gates:
//             with L_k = U_m .. U_{k+1}, R_k = U_{k-1} .. U_1 the gradient of U_k (embedded) is
//             E_k = L_k^T G_W R_k^T, updated from gate to gate by E_{k+1} = conj(U_{k+1}) E_k U_k^T, and the 4 x 4
//             (2 x 2, diagonal) reference gradient is its partial trace over the other window qubits --
//             the reverse pass of src/circuit.rs:320-392 carried out on 64 x 64 matrices instead of 2^n vectors.
//
// Every other window (NonU gates, registers below 2^14) takes the FP32-pipe tile kernels.
#pragma once
#include <atomic>
#include <thread>

#include "tc_block.cuh"
#include "tc_rev.cuh"
#include "tc_host.hpp"
#include "tile_kernels.cuh"

#ifndef QDC_F64

struct TcState {
  bool attr_set = false;
  uint32_t* d_images = nullptr;   // [slots][kImageW / 4]
  uint32_t* h_images = nullptr;   // pinned mirror
  size_t image_slots = 0, image_used = 0;
  double* d_grads = nullptr;      // [slots][128 * 128]
  size_t grad_slots = 0;
  float* d_partials = nullptr;    // [sm_count][128 * 128]
  int partial_ctas = 0;
  int* d_error = nullptr;
  std::vector<TcPass> passes;                 // reverse passes of the current backward() with a live adjoint
  std::vector<std::vector<zc>> grad_of_inst;  // finished gradients (reference order), indexed by instruction
  void release() {
    if (d_images) cudaFree(d_images);
    if (h_images) cudaFreeHost(h_images);
    if (d_grads) cudaFree(d_grads);
    if (d_partials) cudaFree(d_partials);
    if (d_error) cudaFree(d_error);
    *this = TcState();
  }
};

inline void Circuit::release_tc() {
  if (tc_) {
    tc_->release();
    delete tc_;
    tc_ = nullptr;
  }
}

// Is this TILE step a tensor-core block?  (decided from the plan and the instruction kinds only)
inline bool Circuit::tc_step_ok(const qdc::Step& t) const {
  if (!tc_active() || t.tb_count > tcb::kBlockQubits || t.count < 2) return false;
  for (int k = 0; k < t.count; k++)
    if (kind_is_nonu(insts_[plan_.tile_steps[t.first + k].inst].kind)) return false;
  return true;
}

inline const char* Circuit::tc_ensure(size_t image_slots, size_t grad_slots) {
  if (!tc_) tc_ = new TcState();
  TcState& tc = *tc_;
  DeviceInfo di;
  QDC_TRY(qdc_device_info(&di));
  if (!tc.attr_set) {
    QDC_CUDA(cudaFuncSetAttribute(tcb::k_tc_block_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::kSmemBytes));
    QDC_CUDA(cudaFuncSetAttribute(tcb::k_tc_block_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::kGradSmemBytes));
    QDC_CUDA(cudaFuncSetAttribute(tcb::k_tc_block_rev, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::kRevSmemBytes));
    QDC_CUDA(cudaMalloc((void**)&tc.d_error, sizeof(int)));
    QDC_CUDA(cudaMemset(tc.d_error, 0, sizeof(int)));
    tc.attr_set = true;
  }
  if (image_slots > tc.image_slots) {
    if (tc.d_images) QDC_CUDA(cudaFree(tc.d_images));
    if (tc.h_images) QDC_CUDA(cudaFreeHost(tc.h_images));
    QDC_CUDA(cudaMalloc((void**)&tc.d_images, image_slots * tcb::kImageW));
    QDC_CUDA(cudaMallocHost((void**)&tc.h_images, image_slots * tcb::kImageW));
    tc.image_slots = image_slots;
  }
  if (grad_slots > tc.grad_slots) {
    if (tc.d_grads) QDC_CUDA(cudaFree(tc.d_grads));
    QDC_CUDA(cudaMalloc((void**)&tc.d_grads, grad_slots * tcb::kDim * tcb::kDim * sizeof(double)));
    tc.grad_slots = grad_slots;
  }
  if (grad_slots > 0 && tc.partial_ctas < di.sm_count) {
    if (tc.d_partials) QDC_CUDA(cudaFree(tc.d_partials));
    QDC_CUDA(cudaMalloc((void**)&tc.d_partials, (size_t)di.sm_count * tcb::kDim * tcb::kDim * sizeof(float)));
    tc.partial_ctas = di.sm_count;
  }
  return nullptr;
}

// Start of a forward sweep / backward call: size the image ring and the gradient slots for the plan.
inline const char* Circuit::tc_begin(bool backward) {
  if (tc_) {
    tc_->image_used = 0;
    tc_->passes.clear();
    tc_->grad_of_inst.clear();
  }
  if (!tc_active()) return nullptr;
  size_t n = 0;
  for (const qdc::Step& st : plan_.steps)
    if (st.type == qdc::ST_TILE && tc_step_ok(st)) n++;
  if (n == 0) return nullptr;
  QDC_TRY(tc_ensure(backward ? 2 * n : n, backward ? n : 0));
  tc_->image_used = 0;
  return nullptr;
}

inline const std::vector<zc>* Circuit::tc_gradient(size_t inst) const {
  if (!tc_ || inst >= tc_->grad_of_inst.size() || tc_->grad_of_inst[inst].empty()) return nullptr;
  return &tc_->grad_of_inst[inst];
}

// Geometry of the block of a TILE step and its gates with their kernel index bits.
inline const char* Circuit::tc_prepare(const qdc::Step& t, const std::vector<const cplx_t*>& gp, tcb::Params* geo,
                                       TcPass* pass) {
  int block[6], nb = 0;
  for (int k = 0; k < t.tb_count; k++) block[nb++] = plan_.tile_bits[t.tb_first + k];
  for (int q = 3; nb < 6 && q < n_loc_; q++) {   // pad with idle positions >= 3 (W is the identity on them)
    bool used = false;
    for (int k = 0; k < nb; k++) used |= block[k] == q;
    if (!used) block[nb++] = q;
  }
  if (nb < 6) return qdc_errf("register too small for a tensor-core block.");
  int wbit[6];
  const char* e = tcb::make_params(block, n_loc_, geo, wbit);
  if (e) return qdc_errf("%s", e);
  int kbit_of_pos[64];
  for (int q = 0; q < 64; q++) kbit_of_pos[q] = -1;
  for (int k = 0; k < 6; k++) kbit_of_pos[block[wbit[k]]] = k;
  pass->gates.clear();
  for (int k = 0; k < t.count; k++) {
    const qdc::Step& st = plan_.tile_steps[t.first + k];
    const int kind = insts_[st.inst].kind;
    TcGate g;
    g.inst = st.inst;
    g.diag = kind_is_diag(kind);
    g.nq = kind_is_q1(kind) ? 1 : 2;
    g.b2 = kbit_of_pos[st.p2];
    g.b1 = g.nq == 2 ? kbit_of_pos[st.p1] : -1;
    if (g.b2 < 0 || (g.nq == 2 && g.b1 < 0)) return qdc_errf("internal: gate outside its tensor-core block.");
    const cplx_t* src = gp[st.inst];
    for (int i = 0; i < 16; i++) g.m[i] = zc(0, 0);
    if (g.diag) {
      for (int i = 0; i < 4; i++) g.m[5 * i] = zc((double)src[i].x, (double)src[i].y);
    } else {
      const int len = g.nq == 1 ? 4 : 16;
      for (int i = 0; i < len; i++) g.m[i] = zc((double)src[i].x, (double)src[i].y);
    }
    pass->gates.push_back(g);
  }
  return nullptr;
}

// form: 0 = W, 1 = W^dagger, 2 = W^T.  launch = false: only stage the image (geo_out->w_image) for another kernel.
inline const char* Circuit::tc_launch_block(cplx_t* buf, const tcb::Params& geo_in, const Mat64& w, int form, bool launch,
                                            tcb::Params* geo_out) {
  TcState& tc = *tc_;
  if (tc.image_used >= tc.image_slots) return qdc_errf("internal: tensor-core image ring exhausted.");
  std::vector<double> flat(64 * 64 * 2);
  for (int i = 0; i < 64; i++)
    for (int j = 0; j < 64; j++) {
      const zc v = form == 0 ? w[i * 64 + j] : (form == 1 ? std::conj(w[j * 64 + i]) : w[j * 64 + i]);
      flat[2 * (i * 64 + j)] = v.real();
      flat[2 * (i * 64 + j) + 1] = v.imag();
    }
  const std::vector<uint32_t> img = tcb::make_w_image(flat.data());
  uint32_t* h = tc.h_images + tc.image_used * (tcb::kImageW / 4);
  uint32_t* d = tc.d_images + tc.image_used * (tcb::kImageW / 4);
  tc.image_used++;
  memcpy(h, img.data(), tcb::kImageW);
  QDC_CUDA(cudaMemcpyAsync(d, h, tcb::kImageW, cudaMemcpyHostToDevice, stream_));
  tcb::Params geo = geo_in;
  geo.w_image = d;
  geo.error_flag = tc.d_error;
  geo.products = opt_tc_products_;
  if (geo_out) *geo_out = geo;
  if (!launch) return nullptr;
  DeviceInfo di;
  QDC_TRY(qdc_device_info(&di));
  const int grid = (int)std::min<uint64_t>(geo.ntiles, (uint64_t)di.sm_count);
  tcb::k_tc_block_fwd<<<grid, tcb::kThreads, tcb::kSmemBytes, stream_>>>((float2*)buf, geo);
  QDC_CUDA(cudaGetLastError());
  return nullptr;
}

// forward (uncompute = false) or un-compute without an adjoint (uncompute = true) of one block
inline const char* Circuit::run_tc_forward(const qdc::Step& t, const std::vector<const cplx_t*>& gp, bool uncompute) {
  tcb::Params geo;
  TcPass pass;
  QDC_TRY(tc_prepare(t, gp, &geo, &pass));
  Mat64 w;
  tc_block_matrix(pass, w);
  cudaEvent_t pa = nullptr;
  if (prof_.on) pa = prof_.begin(stream_);
  QDC_TRY(tc_launch_block(state_, geo, w, uncompute ? 1 : 0, true));
  if (prof_.on) prof_.end(stream_, uncompute ? CAT_UNCOMPUTE : CAT_TC_FWD, pa, 2ull * t.count * bytes());
  stats_.kernel_launches += 1;
  stats_.hbm_passes += 1;
  stats_.algorithmic_bytes += 2ull * t.count * bytes();
  return nullptr;
}

// reverse step of one block with a live adjoint
inline const char* Circuit::run_tc_backward(const qdc::Step& t, const std::vector<const cplx_t*>& gp) {
  TcState& tc = *tc_;
  tcb::Params geo;
  tc.passes.emplace_back();
  TcPass& pass = tc.passes.back();
  QDC_TRY(tc_prepare(t, gp, &geo, &pass));
  pass.grad_slot = (int)tc.passes.size() - 1;
  if ((size_t)pass.grad_slot >= tc.grad_slots) return qdc_errf("internal: tensor-core gradient slots exhausted.");
  Mat64 w;
  tc_block_matrix(pass, w);
  cudaEvent_t pa = nullptr;
  if (prof_.on) pa = prof_.begin(stream_);
  DeviceInfo di;
  QDC_TRY(qdc_device_info(&di));
  const int grid = (int)std::min<uint64_t>(geo.ntiles, (uint64_t)di.sm_count);
  double* slot = tc.d_grads + (size_t)pass.grad_slot * tcb::kDim * tcb::kDim;
  QDC_CUDA(cudaMemsetAsync(tc.d_partials, 0, (size_t)grid * tcb::kDim * tcb::kDim * sizeof(float), stream_));
  pass.from_h = opt_tc_rev_ != 0;
  if (opt_tc_rev_) {
    // one sweep: state <- W^dagger state, adjoint <- W^T adjoint, H += adjoint (x) state
    tcb::RevParams rp;
    QDC_TRY(tc_launch_block(nullptr, geo, w, 1, false, &rp.geo));
    rp.partials = tc.d_partials;
    rp.h_products = 6;
    tcb::k_tc_block_rev<<<grid, tcb::kRevThreads, tcb::kRevSmemBytes, stream_>>>((float2*)state_, (float2*)bwd_, rp);
    QDC_CUDA(cudaGetLastError());
    tcb::k_tc_grad_reduce<<<(tcb::kDim * tcb::kDim + 255) / 256, 256, 0, stream_>>>(tc.d_partials, grid, slot, 0);
    QDC_CUDA(cudaGetLastError());
    stats_.kernel_launches += 2;
    stats_.hbm_passes += 2;
  } else {
    QDC_TRY(tc_launch_block(state_, geo, w, 1, true));                 // state <- W^dagger state
    tcb::GradParams gpar;
    gpar.geo = geo;
    gpar.geo.error_flag = tc.d_error;
    gpar.geo.w_image = nullptr;
    gpar.geo.products = 6;
    gpar.partials = tc.d_partials;
    tcb::k_tc_block_grad<<<grid, tcb::kThreads, tcb::kGradSmemBytes, stream_>>>((const float2*)state_, (const float2*)bwd_, gpar);
    QDC_CUDA(cudaGetLastError());
    tcb::k_tc_grad_reduce<<<(tcb::kDim * tcb::kDim + 255) / 256, 256, 0, stream_>>>(tc.d_partials, grid, slot, 0);
    QDC_CUDA(cudaGetLastError());
    QDC_TRY(tc_launch_block(bwd_, geo, w, 2, true));                   // adjoint <- W^T adjoint
    stats_.kernel_launches += 4;
    stats_.hbm_passes += 3;
  }
  if (prof_.on) prof_.end(stream_, CAT_TC_BWD, pa, 4ull * t.count * bytes());
  stats_.algorithmic_bytes += 4ull * t.count * bytes();
  return nullptr;
}

// After the stream has been synchronised: block gradients -> gate gradients (reference order) by the chain rule.
inline const char* Circuit::tc_finish_backward() {
  if (!tc_) return nullptr;
  TcState& tc = *tc_;
  tc.grad_of_inst.assign(insts_.size(), std::vector<zc>());
  const size_t np = tc.passes.size();
  if (np == 0) return nullptr;
  int herr = 0;
  QDC_CUDA(cudaMemcpy(&herr, tc.d_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr) return qdc_errf("tensor-core block kernel: pipeline watchdog fired.");
  std::vector<double> P(np * tcb::kDim * tcb::kDim);
  QDC_CUDA(cudaMemcpy(P.data(), tc.d_grads, P.size() * sizeof(double), cudaMemcpyDeviceToHost));
  std::atomic<size_t> next(0);
  auto work = [&]() {
    for (;;) {
      const size_t pi = next.fetch_add(1);
      if (pi >= np) return;
      const TcPass& pass = tc.passes[pi];
      const double* pp = &P[(size_t)pass.grad_slot * tcb::kDim * tcb::kDim];
      auto want = [&](int k) { return kind_is_var(insts_[pass.gates[k].inst].kind); };
      auto emit = [&](int k, const zc* v, int count) { tc.grad_of_inst[pass.gates[k].inst].assign(v, v + count); };
      if (pass.from_h) tc_chain_rule_from_h(pass, pp, want, emit);
      else tc_chain_rule(pass, pp, want, emit);
    }
  };
  unsigned nt = std::thread::hardware_concurrency();
  nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
  if (np < 4) nt = 1;
  std::vector<std::thread> pool;
  for (unsigned k = 1; k < nt; k++) pool.emplace_back(work);
  work();
  for (std::thread& th : pool) th.join();
  return nullptr;
}

#endif  // !QDC_F64
