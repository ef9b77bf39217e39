/* oracle/cpu_ref.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the srvk/eesen BiLSTM-parallel + CTC training
 * hot path.  It is the checker for the CUDA product in eesen_b200/csrc; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference
 * leg may load it.  The product never links or calls it.
 *
 * Pinning status: the BiLSTM/affine/softmax/SGD parts are pinned against the
 * UNMODIFIED reference compiled from /root/reference (oracle/_ref/ref_dump_cpu,
 * tests/test_oracle_vs_reference.py + committed fixtures tests/golden/).  The
 * CTC part has NO CPU implementation in the reference (cuda-matrix.cc:862-864,
 * 895-897, 928-930, 962-964, 995-997, 1031-1033 "not implemented for CPU yet")
 * and the reference ships no tests for this path: it is
 * pinned against the reference's own CUDA kernels run on the B200 box
 * (oracle/_ref/ref_dump_gpu, fixtures generated there and committed under
 * tests/golden/) and cross-checked against torch.nn.functional.ctc_loss (fp64).
 *
 * Compiled twice: -DREAL=float (liboracle_f32.so, the "port" CPU baseline and
 * like-for-like fp32 checker) and -DREAL=double (liboracle_f64.so, the arbiter).
 *
 * Layout conventions (reference: src/netbin/train-ctc-parallel.cc:186-193):
 *   packed minibatch row r = t*S + s (frame t of utterance s), zero padded to T.
 *   BiLSTM state buffers have T+2 time slots (slot 0 and T+1 are zero boundary
 *   rows), 7 column blocks of C: g,i,f,o,c,h,m (bilstm-parallel-layer.h:99-107).
 *   BiLSTM params, 12 tensors in the order of BiLstm::WriteData
 *   (bilstm-layer.h:478-492): wx_fw[4C x I], wm_fw[4C x C], b_fw[4C], pi_fw[C],
 *   pf_fw[C], po_fw[C], then the same six for bw.
 */
#ifndef EESEN_B200_ORACLE_CPU_REF_H_
#define EESEN_B200_ORACLE_CPU_REF_H_

#ifndef REAL
#define REAL float
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* BiLstmParallel::PropagateFnc (bilstm-parallel-layer.h:379-420, vanilla passes :97-206).
 * buf_fw/buf_bw: [(T+2)*S x 7C] (fully overwritten), out: [T*S x 2C]. */
void oracle_bilstm_forward(int T, int S, int I, int C, const int *len, const REAL *x,
                           const REAL *const *params, REAL *buf_fw, REAL *buf_bw, REAL *out);

/* BiLstmParallel::BackpropagateFnc (:881-913, vanilla passes :422-602).
 * dbuf_fw/dbuf_bw: [(T+2)*S x 7C] scratch (overwritten), in_diff: [T*S x I],
 * corr: 12 gradient accumulators, updated as corr = grad + momentum*corr. */
void oracle_bilstm_backward(int T, int S, int I, int C, const REAL *x,
                            const REAL *const *params, const REAL *buf_fw, const REAL *buf_bw,
                            const REAL *out_diff, REAL *dbuf_fw, REAL *dbuf_bw, REAL *in_diff,
                            REAL *const *corr, REAL momentum);

/* The recurrent-dropout passes of BiLstmParallel (bilstm-parallel-layer.h:209-377, :604-879).  drop: 1 =
 * no-mem-loss dropout (mask on g*i), 2 = RNNdrop (mask on the whole cell).  rmask_*: scaled masks of one
 * direction, [T*S x C] if per_step (row (t-1)*S+s) else [S x C].  Forward dropout (mask on the layer OUTPUT,
 * :409-416, and on out_diff, :891-895) is an elementwise product applied by the caller. */
void oracle_bilstm_forward_drop(int T, int S, int I, int C, const int *len, const REAL *x, const REAL *const *params,
                                REAL *buf_fw, REAL *buf_bw, REAL *out, int drop, const REAL *rmask_fw,
                                const REAL *rmask_bw, int per_step);
void oracle_bilstm_backward_drop(int T, int S, int I, int C, const REAL *x, const REAL *const *params,
                                 const REAL *buf_fw, const REAL *buf_bw, const REAL *out_diff, REAL *dbuf_fw,
                                 REAL *dbuf_bw, REAL *in_diff, REAL *const *corr, REAL momentum, int drop,
                                 const REAL *rmask_fw, const REAL *rmask_bw, int per_step);

/* LstmParallel::PropagateFnc / BackpropagateFnc (lstm-parallel-layer.h:47-113, :115-213): 6 params in the
 * order of Lstm::WriteData (lstm-layer.h:147-172): wx, wm, bias, pi, pf, po.  buf/dbuf: [(T+2)*S x 7C]. */
void oracle_lstm_forward(int T, int S, int I, int C, const REAL *x, const REAL *const *params, REAL *buf, REAL *out);
void oracle_lstm_backward(int T, int S, int I, int C, const REAL *x, const REAL *const *params, const REAL *buf,
                          const REAL *out_diff, REAL *dbuf, REAL *in_diff, REAL *const *corr, REAL momentum);

/* AffineTransform::PropagateFnc (affine-trans-layer.h:161-166): out = in*W^T + b */
void oracle_affine_forward(int N, int D, int K, const REAL *in, const REAL *W, const REAL *b, REAL *out);
/* AffineTransform::BackpropagateFnc (:168-172): in_diff = out_diff * W */
void oracle_affine_backward(int N, int D, int K, const REAL *out_diff, const REAL *W, REAL *in_diff);
/* AffineTransform::Update gradient part (:182-183): Wc = diff^T*in + mu*Wc ; bc = colsum(diff) + mu*bc */
void oracle_affine_grad(int N, int D, int K, const REAL *in, const REAL *diff, REAL *Wc, REAL *bc, REAL momentum);

/* Softmax::PropagateFnc -> ApplySoftMaxPerRow (softmax-layer.h:46; kernel cuda-kernels.cu:744-808) */
void oracle_softmax(int N, int K, const REAL *in, REAL *out);

/* Ctc::EvalParallel (ctc-loss.cc:101-194) with the MSeq kernels
 * (cuda-kernels.cu:1369-1408, 1484-1544, 1605-1627) and log-math ctc-utils.h:29-96.
 * y: softmax probabilities [T*S x K]; labels: concatenated, lab_len[s] each (1-based ids);
 * outputs: pzx[S], diff[T*S x K]; alpha/beta ([T*S x Lp], Lp = 2*max_lab+1) may be NULL. */
void oracle_ctc_eval_parallel(int T, int S, int K, const int *len, const int *labels, const int *lab_len,
                              const REAL *y, REAL *pzx, REAL *diff, REAL *alpha, REAL *beta);

/* Clip (ApplyFloor/ApplyCeiling, bilstm-layer.h:848-862; affine :186-189) + SGD step
 * (bilstm-layer.h:865-883; affine :191-195): corr = clamp(corr, +-max_grad) in place if max_grad>0;
 * w -= lr*corr. */
void oracle_sgd_update(long n, REAL *w, REAL *corr, REAL lr, REAL max_grad);

/* Adagrad (mode 1) / RMSProp (mode 2) update of one tensor; see cpu_ref.c */
void oracle_ada_update(long n, REAL *w, REAL *corr, REAL *accu, REAL lr, REAL max_grad, REAL eps, REAL rho,
                       REAL one_minus_rho, int mode);

/* FindRowMaxId (cuda-matrix.cc:1038-1095): first index of the row maximum. */
void oracle_row_argmax(int N, int K, const REAL *y, int *idx);

int oracle_real_size(void);

#ifdef __cplusplus
}
#endif

/* 1: round both operands of every dense (input-side / affine) product to bf16; 0: exact (default) */
void oracle_set_dense_rounding(int mode);

#endif
