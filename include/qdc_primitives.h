/* qdc_primitives.h -- the drop-in C ABI of the differentiable statevector
 * hot path: the 18 `extern "C"` symbols the reference's Rust host code binds
 * in /root/reference/src/primitives_bind.rs:15-119 and defines in
 * /root/reference/src/primitives.cu.  A library built from this repo
 * (libqdc_b200_f32.so / libqdc_b200_f64.so) exports exactly these names with
 * exactly these signatures, so `build.rs` can link it in place of
 * libprimitives.a.
 *
 * One precision per library, as in the reference (cargo feature `f64` ->
 * -DF64, src/primitives.cu:11-29): compile the consumer with -DQDC_F64 when
 * linking the f64 build.
 *
 * Conventions (identical to the reference):
 *  - `qdc_complex` is an interleaved (re, im) pair == cuFloatComplex /
 *    cuDoubleComplex == num_complex::Complex<f32|f64> == NumPy complex64/128.
 *  - A state is 2^qubits_number amplitudes in DEVICE memory; qubit k is bit k
 *    of the linear index (qubit 0 innermost).
 *  - Gate pointers are HOST memory, flat row-major (4 or 16 entries), fully
 *    consumed before the call returns.
 *  - Density / gradient outputs are HOST memory and are ACCUMULATED INTO
 *    (`out[i] += result[i]`, src/primitives.cu:281-288); the caller zeroes them.
 *  - Fallible functions return NULL on success, otherwise a heap-allocated
 *    message the caller may print and need not free (src/primitives.cu:32-49).
 *  - Everything runs on the legacy default stream; calls are not thread-safe
 *    with respect to one state, as in the reference (README.md:13).
 *  - Unlike the reference (`1 << n` with int, src/primitives.cu:147 etc.),
 *    all sizes are 64-bit: qubits_number up to 34 is supported.
 */
#ifndef QDC_PRIMITIVES_H
#define QDC_PRIMITIVES_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef QDC_F64
typedef struct { double re, im; } qdc_complex;
#else
typedef struct { float re, im; } qdc_complex;
#endif

/* src/primitives.cu:189-199  (primitives_bind.rs:16-19) */
void set2standard(qdc_complex* state, size_t qubits_number);
/* src/primitives.cu:141-150  (primitives_bind.rs:20-23) */
const char* get_state(qdc_complex** state, size_t qubits_number);
/* src/primitives.cu:166-173  (primitives_bind.rs:24) */
const char* drop_state(qdc_complex* state);
/* src/primitives.cu:153-163  (primitives_bind.rs:25-29) */
const char* copy_to_host(const qdc_complex* state, qdc_complex* host_state, size_t qubits_number);
/* src/primitives.cu:496-510  (primitives_bind.rs:64-68) */
const char* set_from_host(qdc_complex* device_state, const qdc_complex* host_state, size_t qubits_number);

/* src/primitives.cu:534-545  (primitives_bind.rs:30-35): psi'[p] = sum_q g[2p+q] psi[q] at bit pos */
const char* q1gate(qdc_complex* state, const qdc_complex* gate, size_t pos, size_t qubits_number);
/* src/primitives.cu:547-570  (primitives_bind.rs:36-41): applies gate^-1 */
const char* q1gate_inv(qdc_complex* state, const qdc_complex* gate, size_t pos, size_t qubits_number);
/* src/primitives.cu:608-620  (primitives_bind.rs:42-48): gate[8 q2 + 4 q1 + 2 p2 + p1] */
const char* q2gate(qdc_complex* state, const qdc_complex* gate, size_t pos2, size_t pos1, size_t qubits_number);
/* src/primitives.cu:622-646  (primitives_bind.rs:49-55) */
const char* q2gate_inv(qdc_complex* state, const qdc_complex* gate, size_t pos2, size_t pos1, size_t qubits_number);
/* src/primitives.cu:674-686  (primitives_bind.rs:56-62): psi[p2,p1] *= gate[2 p2 + p1] */
const char* q2gate_diag(qdc_complex* state, const qdc_complex* gate, size_t pos2, size_t pos1, size_t qubits_number);

/* src/primitives.cu:741-776  (primitives_bind.rs:69-74): density[2p+q] += sum psi[p] conj psi[q] */
const char* get_q1density(const qdc_complex* state, qdc_complex* density, size_t pos, size_t qubits_number);
/* src/primitives.cu:839-876  (primitives_bind.rs:75-81): density[8p2+4p1+2q2+q1] += ... */
const char* get_q2density(const qdc_complex* state, qdc_complex* density, size_t pos2, size_t pos1, size_t qubits_number);

/* src/primitives.cu:255-292  (primitives_bind.rs:82-88): grad[2p+q] += sum bwd[p] fwd[q] */
const char* q1grad(const qdc_complex* fwd, const qdc_complex* bwd, qdc_complex* grad, size_t pos, size_t qubits_number);
/* src/primitives.cu:356-395  (primitives_bind.rs:89-96) */
const char* q2grad(const qdc_complex* fwd, const qdc_complex* bwd, qdc_complex* grad, size_t pos2, size_t pos1, size_t qubits_number);
/* src/primitives.cu:454-493  (primitives_bind.rs:97-104) */
const char* q2grad_diag(const qdc_complex* fwd, const qdc_complex* bwd, qdc_complex* grad, size_t pos2, size_t pos1, size_t qubits_number);

/* src/primitives.cu:917-929  (primitives_bind.rs:105-109): dst = 2 conj(src) */
void conj_and_double(const qdc_complex* src, qdc_complex* dst, size_t qubits_number);
/* src/primitives.cu:941-953  (primitives_bind.rs:110-114): dst += src */
void add(const qdc_complex* src, qdc_complex* dst, size_t qubits_number);
/* src/primitives.cu:889-901  (primitives_bind.rs:115-119): dst = src */
void copy(const qdc_complex* src, qdc_complex* dst, size_t qubits_number);

#ifdef __cplusplus
}
#endif
#endif /* QDC_PRIMITIVES_H */
