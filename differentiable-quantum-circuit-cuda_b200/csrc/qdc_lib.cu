// libqdc_b200_{f32,f64}.so -- single translation unit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 [-DQDC_F64] -shared ...
// Exports: the 18 legacy symbols (include/qdc_primitives.h) and the circuit
// ABI (include/qdc_circuit.h).  Everything else has hidden visibility.
#include "primitives_abi.cuh"
#include "circuit.cuh"

struct qdc_circuit {
  Circuit impl;
  explicit qdc_circuit(int n) : impl(n) {}
};

QDC_EXPORT const char* qdc_precision(void) {
#ifdef QDC_F64
  return "f64";
#else
  return "f32";
#endif
}

QDC_EXPORT int qdc_abi_version(void) { return 1; }

QDC_EXPORT const char* qdc_circuit_new(qdc_circuit** out, size_t qubits_number) {
  if (qubits_number < 1 || qubits_number > 40) return qdc_errf("qubits_number out of range.");
  qdc_circuit* c = new qdc_circuit((int)qubits_number);
  const char* e = c->impl.ensure_state();
  if (e) {
    delete c;
    return e;
  }
  *out = c;
  return nullptr;
}

QDC_EXPORT const char* qdc_circuit_free(qdc_circuit* c) {
  delete c;
  return nullptr;
}

QDC_EXPORT const char* qdc_circuit_set_state_from_host(qdc_circuit* c, const cplx_t* host_state, size_t len) {
  return c->impl.set_state_from_host(host_state, len);
}

QDC_EXPORT const char* qdc_circuit_add(qdc_circuit* c, int kind, size_t pos2, size_t pos1) {
  return c->impl.add(kind, pos2, pos1);
}

QDC_EXPORT size_t qdc_circuit_count(const qdc_circuit* c, int what) { return c->impl.count(what); }

QDC_EXPORT const char* qdc_circuit_run(qdc_circuit* c, const cplx_t* cgates, const uint32_t* clens, size_t nc,
                                       const cplx_t* vgates, const uint32_t* vlens, size_t nv, cplx_t* out,
                                       size_t cap, size_t* out_len) {
  return c->impl.sweep(GateList{cgates, clens, nc}, GateList{vgates, vlens, nv}, true, out, cap, out_len);
}

QDC_EXPORT const char* qdc_circuit_forward(qdc_circuit* c, const cplx_t* cgates, const uint32_t* clens,
                                           size_t nc, const cplx_t* vgates, const uint32_t* vlens, size_t nv,
                                           cplx_t* out, size_t cap, size_t* out_len) {
  return c->impl.sweep(GateList{cgates, clens, nc}, GateList{vgates, vlens, nv}, false, out, cap, out_len);
}

QDC_EXPORT const char* qdc_circuit_backward(qdc_circuit* c, const cplx_t* dgrads, const uint32_t* dlens,
                                            size_t nd, const cplx_t* cgates, const uint32_t* clens, size_t nc,
                                            const cplx_t* vgates, const uint32_t* vlens, size_t nv, cplx_t* out,
                                            size_t cap, size_t* out_len) {
  return c->impl.backward(GateList{dgrads, dlens, nd}, GateList{cgates, clens, nc},
                          GateList{vgates, vlens, nv}, out, cap, out_len);
}

QDC_EXPORT const char* qdc_circuit_copy_state_to_host(qdc_circuit* c, cplx_t* host_state) {
  QDC_TRY(c->impl.ensure_state());
  QDC_CUDA(cudaStreamSynchronize(c->impl.stream_));
  QDC_CUDA(cudaMemcpy(host_state, c->impl.state_, c->impl.bytes(), cudaMemcpyDeviceToHost));
  return nullptr;
}

QDC_EXPORT const char* qdc_circuit_state_device_ptr(qdc_circuit* c, void** device_ptr) {
  QDC_TRY(c->impl.ensure_state());
  *device_ptr = c->impl.state_;
  return nullptr;
}

QDC_EXPORT const char* qdc_circuit_set_stream(qdc_circuit* c, void* cuda_stream) {
  c->impl.stream_ = (cudaStream_t)cuda_stream;
  return nullptr;
}

QDC_EXPORT const char* qdc_circuit_set_option(qdc_circuit* c, const char* key, long value) {
  if (strcmp(key, "fuse") == 0) {
    c->impl.opt_fuse_ = (int)value;
    return nullptr;
  }
  if (strcmp(key, "profile") == 0) {
    c->impl.prof_.on = value != 0;
    return nullptr;
  }
  return qdc_errf("Unknown option \"%s\".", key);
}

typedef struct {
  uint64_t kernel_launches, hbm_passes, algorithmic_bytes;
} qdc_stats;

QDC_EXPORT const char* qdc_circuit_last_stats(const qdc_circuit* c, qdc_stats* out) {
  out->kernel_launches = c->impl.stats_.kernel_launches;
  out->hbm_passes = c->impl.stats_.hbm_passes;
  out->algorithmic_bytes = c->impl.stats_.algorithmic_bytes;
  return nullptr;
}

typedef struct {
  uint64_t launches;
  double ms;
  uint64_t algorithmic_bytes;
} qdc_profile_entry;

QDC_EXPORT int qdc_profile_categories(void) { return CAT_COUNT; }

QDC_EXPORT const char* qdc_profile_category_name(int cat) {
  return (cat >= 0 && cat < CAT_COUNT) ? kProfCatNames[cat] : "";
}

QDC_EXPORT const char* qdc_circuit_last_profile(const qdc_circuit* c, int cat, qdc_profile_entry* out) {
  if (cat < 0 || cat >= CAT_COUNT) return qdc_errf("Unknown profile category %d.", cat);
  out->launches = c->impl.prof_.cats[cat].launches;
  out->ms = c->impl.prof_.cats[cat].ms;
  out->algorithmic_bytes = c->impl.prof_.cats[cat].alg_bytes;
  return nullptr;
}

QDC_EXPORT const char* qdc_reverse_step(cplx_t* fwd, cplx_t* bwd, const cplx_t* gate, cplx_t* grad, int kind,
                                        int non_unitary, size_t pos2, size_t pos1, size_t n) {
  const int inv_form = non_unitary ? FORM_INV : FORM_CONJ_TR;
  double* dst = nullptr;
  if (grad) {
    QDC_TRY(ws_ensure(g_legacy_ws, 1));
    dst = g_legacy_ws.red_out;
  }
  if (kind_is_q1(kind)) {
    QDC_TRY(eng_rev_q1(0, g_legacy_ws, fwd, bwd, gate, inv_form, (int)pos2, (int)n, dst));
    return grad ? fetch_and_accumulate(8, false, false, grad) : nullptr;
  }
  if (kind_is_q2dense(kind)) {
    QDC_TRY(eng_rev_q2(0, g_legacy_ws, fwd, bwd, gate, inv_form, (int)pos2, (int)pos1, (int)n, dst));
    return grad ? fetch_and_accumulate(32, true, pos2 < pos1, grad) : nullptr;
  }
  if (kind_is_diag(kind)) {
    QDC_TRY(eng_rev_diag(0, g_legacy_ws, fwd, bwd, gate, (int)pos2, (int)pos1, (int)n, dst));
    return grad ? fetch_and_accumulate(8, false, false, grad) : nullptr;
  }
  return qdc_errf("qdc_reverse_step: kind %d is not a gate.", kind);
}

QDC_EXPORT const char* qdc_density_seed(const cplx_t* fwd, cplx_t* bwd, const cplx_t* dens_grad, int kind,
                                        int accumulate, size_t pos2, size_t pos1, size_t n) {
  if (kind_is_q1(kind) || kind_is_q1_dens(kind))
    return eng_seed_q1(0, g_legacy_ws, fwd, bwd, dens_grad, (int)pos2, (int)n, accumulate != 0);
  return eng_seed_q2(0, g_legacy_ws, fwd, bwd, dens_grad, (int)pos2, (int)pos1, (int)n, accumulate != 0);
}
