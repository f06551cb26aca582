// The 18 legacy C symbols (include/qdc_primitives.h) on top of the sm_100a
// streaming kernels.  Same names, signatures, accumulate-into semantics and
// error convention as /root/reference/src/primitives.cu; different engine.
#pragma once
#include "engine.cuh"

static Workspace g_legacy_ws;  // legacy entry points: default stream, one shared workspace

#define QDC_EXPORT extern "C" __attribute__((visibility("default")))

// Synchronous read-back of a reduction result and `+=` into the caller's
// host buffer (src/primitives.cu:281-288 and siblings).
static const char* fetch_and_accumulate(int nred, bool is_q2, bool swap, cplx_t* host_out) {
  double h[32];
  QDC_CUDA(cudaMemcpy(h, g_legacy_ws.red_out, nred * sizeof(double), cudaMemcpyDeviceToHost));
  if (is_q2) {
    zc m[16];
    unpermute_q2(h, swap, m);
    for (int i = 0; i < 16; i++) {
      host_out[i].x += (real_t)m[i].real();
      host_out[i].y += (real_t)m[i].imag();
    }
  } else {
    for (int i = 0; i < nred / 2; i++) {
      host_out[i].x += (real_t)h[2 * i];
      host_out[i].y += (real_t)h[2 * i + 1];
    }
  }
  return nullptr;
}

QDC_EXPORT const char* get_state(cplx_t** state, size_t qubits_number) {
  QDC_CUDA(cudaMalloc((void**)state, sizeof(cplx_t) << qubits_number));
  return nullptr;
}

QDC_EXPORT const char* drop_state(cplx_t* state) {
  QDC_CUDA(cudaFree(state));
  return nullptr;
}

QDC_EXPORT const char* copy_to_host(const cplx_t* state, cplx_t* host_state, size_t qubits_number) {
  QDC_CUDA(cudaMemcpy(host_state, state, sizeof(cplx_t) << qubits_number, cudaMemcpyDeviceToHost));
  return nullptr;
}

QDC_EXPORT const char* set_from_host(cplx_t* device_state, const cplx_t* host_state, size_t qubits_number) {
  QDC_CUDA(cudaMemcpy(device_state, host_state, sizeof(cplx_t) << qubits_number, cudaMemcpyHostToDevice));
  return nullptr;
}

QDC_EXPORT void set2standard(cplx_t* state, size_t qubits_number) {
  (void)eng_set_standard(0, state, (int)qubits_number);
}

QDC_EXPORT const char* q1gate(cplx_t* state, const cplx_t* gate, size_t pos, size_t qubits_number) {
  return eng_q1gate(0, g_legacy_ws, state, gate, FORM_PLAIN, (int)pos, (int)qubits_number);
}

QDC_EXPORT const char* q1gate_inv(cplx_t* state, const cplx_t* gate, size_t pos, size_t qubits_number) {
  return eng_q1gate(0, g_legacy_ws, state, gate, FORM_INV, (int)pos, (int)qubits_number);
}

QDC_EXPORT const char* q2gate(cplx_t* state, const cplx_t* gate, size_t pos2, size_t pos1,
                              size_t qubits_number) {
  return eng_q2gate(0, g_legacy_ws, state, gate, FORM_PLAIN, (int)pos2, (int)pos1, (int)qubits_number);
}

QDC_EXPORT const char* q2gate_inv(cplx_t* state, const cplx_t* gate, size_t pos2, size_t pos1,
                                  size_t qubits_number) {
  return eng_q2gate(0, g_legacy_ws, state, gate, FORM_INV, (int)pos2, (int)pos1, (int)qubits_number);
}

QDC_EXPORT const char* q2gate_diag(cplx_t* state, const cplx_t* gate, size_t pos2, size_t pos1,
                                   size_t qubits_number) {
  return eng_q2diag(0, g_legacy_ws, state, gate, false, (int)pos2, (int)pos1, (int)qubits_number);
}

QDC_EXPORT const char* get_q1density(const cplx_t* state, cplx_t* density, size_t pos, size_t qubits_number) {
  QDC_TRY(eng_dens_q1(0, g_legacy_ws, state, (int)pos, (int)qubits_number, nullptr));
  return fetch_and_accumulate(8, false, false, density);
}

QDC_EXPORT const char* get_q2density(const cplx_t* state, cplx_t* density, size_t pos2, size_t pos1,
                                     size_t qubits_number) {
  QDC_TRY(eng_dens_q2(0, g_legacy_ws, state, (int)pos2, (int)pos1, (int)qubits_number, nullptr));
  return fetch_and_accumulate(32, true, pos2 < pos1, density);
}

QDC_EXPORT const char* q1grad(const cplx_t* fwd, const cplx_t* bwd, cplx_t* grad, size_t pos,
                              size_t qubits_number) {
  QDC_TRY(eng_grad_q1(0, g_legacy_ws, fwd, bwd, (int)pos, (int)qubits_number, nullptr));
  return fetch_and_accumulate(8, false, false, grad);
}

QDC_EXPORT const char* q2grad(const cplx_t* fwd, const cplx_t* bwd, cplx_t* grad, size_t pos2, size_t pos1,
                              size_t qubits_number) {
  QDC_TRY(eng_grad_q2(0, g_legacy_ws, fwd, bwd, (int)pos2, (int)pos1, (int)qubits_number, nullptr));
  return fetch_and_accumulate(32, true, pos2 < pos1, grad);
}

QDC_EXPORT const char* q2grad_diag(const cplx_t* fwd, const cplx_t* bwd, cplx_t* grad, size_t pos2, size_t pos1,
                                   size_t qubits_number) {
  QDC_TRY(eng_grad_diag(0, g_legacy_ws, fwd, bwd, (int)pos2, (int)pos1, (int)qubits_number, nullptr));
  return fetch_and_accumulate(8, false, false, grad);
}

QDC_EXPORT void conj_and_double(const cplx_t* src, cplx_t* dst, size_t qubits_number) {
  (void)eng_conj_and_double(0, g_legacy_ws, src, dst, (int)qubits_number);
}

QDC_EXPORT void add(const cplx_t* src, cplx_t* dst, size_t qubits_number) {
  (void)eng_add(0, g_legacy_ws, src, dst, (int)qubits_number);
}

QDC_EXPORT void copy(const cplx_t* src, cplx_t* dst, size_t qubits_number) {
  (void)eng_copy(0, src, dst, (int)qubits_number);
}
