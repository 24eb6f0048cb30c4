/*
 * abi_client.c -- a torch-free, Python-free client of libscp_b200.so.
 *
 * Proves the drop-in boundary of include/scp_b200.h: plain C, the CUDA runtime for device memory and a stream, nothing
 * else.  Every result is checked against a scalar double-precision restatement of the reference arithmetic written
 * out below (weighted_sum.py:26-45, kw_branches.py:158-179 + my_vector_quantizer.py:78-82, losses.py:185-245), and the
 * status-code contract of the header (SCP_ERR_INVALID / _UNSUPPORTED / _WORKSPACE, no crash, no device fault) is
 * exercised with bad arguments.  Built and run by tests/test_c_abi_client.py (gcc; run only where a GPU is present).
 *
 *   gcc -std=c11 -O1 -Iinclude -I/usr/local/cuda/include tests/c_client/abi_client.c \
 *       -Lspeechclip_plus_b200 -lscp_b200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,... -o abi_client
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "scp_b200.h"

static int failures = 0;

#define EXPECT(cond, ...)                          \
  do {                                             \
    if (!(cond)) {                                 \
      ++failures;                                  \
      fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
      fprintf(stderr, __VA_ARGS__);                \
      fprintf(stderr, "\n");                       \
    }                                              \
  } while (0)

#define CUDA_OK(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                          \
    }                                                                                   \
  } while (0)

/* xorshift64* -> uniform (0,1) -> Box-Muller; deterministic on every machine */
static uint64_t rng_state = 7122;
static double uniform01(void) {
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0 + 1e-300;
}
static float gauss(void) { return (float)(sqrt(-2.0 * log(uniform01())) * cos(6.283185307179586 * uniform01())); }

static void* dev_alloc(size_t bytes) {
  void* p = NULL;
  CUDA_OK(cudaMalloc(&p, bytes ? bytes : 1));
  CUDA_OK(cudaMemset(p, 0, bytes ? bytes : 1));
  return p;
}
static void* to_dev(const void* host, size_t bytes) {
  void* p = dev_alloc(bytes);
  CUDA_OK(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
  return p;
}
static void to_host(void* host, const void* dev, size_t bytes) { CUDA_OK(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost)); }

/* ------------------------------------------------------------------------------------------------------------ */
/* S1: y = sum_l softmax(w)_l x_l over (T,B,D) storage viewed as (B,T,D); d_w = softmax-Jacobian of sum g*x_l      */
static void check_wsum(cudaStream_t st) {
  enum { L = 5, B = 3, T = 7, D = 16 };
  const size_t n = (size_t)B * T * D;
  float* h_layers[L];
  const void* d_layers[L];
  float w[L], gy[B * T * D];
  for (int l = 0; l < L; ++l) {
    h_layers[l] = (float*)malloc(n * 4); /* storage order (T,B,D) */
    for (size_t i = 0; i < n; ++i) h_layers[l][i] = gauss();
    d_layers[l] = to_dev(h_layers[l], n * 4);
    w[l] = 0.5f * gauss();
  }
  for (size_t i = 0; i < n; ++i) gy[i] = gauss();
  float* d_w = (float*)to_dev(w, sizeof w);
  float* d_gy = (float*)to_dev(gy, sizeof gy);
  float* d_y = (float*)dev_alloc(n * 4);
  float* d_dw = (float*)dev_alloc(L * 4);
  size_t ws_bytes = scp_wsum_bwd_workspace_bytes(L, B, T, D);
  void* d_ws = dev_alloc(ws_bytes);

  /* element (b,t,d) of layer l lives at t*B*D + b*D + d: stride_b = D, stride_t = B*D (speech_encoder_plus.py:596-599) */
  int rc = scp_wsum_fwd(d_layers, L, B, T, D, D, (int64_t)B * D, SCP_F32, d_w, SCP_NORM_NONE, 1e-5f, NULL, d_y, SCP_F32, st);
  EXPECT(rc == SCP_OK, "scp_wsum_fwd -> %d (%s)", rc, scp_last_error_string(rc));
  rc = scp_wsum_bwd(d_layers, L, B, T, D, D, (int64_t)B * D, SCP_F32, d_w, SCP_NORM_NONE, 1e-5f, NULL, d_gy, SCP_F32, d_dw,
                    NULL, d_ws, ws_bytes, st);
  EXPECT(rc == SCP_OK, "scp_wsum_bwd -> %d (%s)", rc, scp_last_error_string(rc));
  CUDA_OK(cudaStreamSynchronize(st));
  float y[B * T * D], dw[L];
  to_host(y, d_y, sizeof y);
  to_host(dw, d_dw, sizeof dw);

  double sm[L], z = 0, dots[L], mix = 0, err_y = 0, max_y = 0;
  for (int l = 0; l < L; ++l) z += exp((double)w[l]);
  for (int l = 0; l < L; ++l) { sm[l] = exp((double)w[l]) / z; dots[l] = 0; }
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t)
      for (int d = 0; d < D; ++d) {
        double acc = 0;
        const size_t out = ((size_t)b * T + t) * D + d, in = ((size_t)t * B + b) * D + d;
        for (int l = 0; l < L; ++l) { acc += sm[l] * h_layers[l][in]; dots[l] += (double)gy[out] * h_layers[l][in]; }
        if (fabs(acc) > max_y) max_y = fabs(acc);
        if (fabs(acc - y[out]) > err_y) err_y = fabs(acc - y[out]);
      }
  EXPECT(err_y / max_y < 1e-5, "wsum forward: rel err %.3g", err_y / max_y);
  for (int l = 0; l < L; ++l) mix += sm[l] * dots[l];
  for (int l = 0; l < L; ++l) {
    const double ref = sm[l] * (dots[l] - mix);
    EXPECT(fabs(ref - dw[l]) < 1e-4 * (fabs(ref) + 1.0), "wsum d_weights[%d]: %.6f vs %.6f", l, dw[l], ref);
  }

  /* status-code contract: bad arguments come back as codes, nothing is launched */
  EXPECT(scp_wsum_fwd(NULL, L, B, T, D, D, B * D, SCP_F32, d_w, 0, 1e-5f, NULL, d_y, SCP_F32, st) == SCP_ERR_INVALID, "null layer array");
  EXPECT(scp_wsum_fwd(d_layers, SCP_MAX_LAYERS + 1, B, T, D, D, B * D, SCP_F32, d_w, 0, 1e-5f, NULL, d_y, SCP_F32, st) == SCP_ERR_INVALID, "L too large");
  EXPECT(scp_wsum_fwd(d_layers, L, 0, T, D, D, B * D, SCP_F32, d_w, 0, 1e-5f, NULL, d_y, SCP_F32, st) == SCP_ERR_INVALID, "B = 0");
  EXPECT(scp_wsum_fwd(d_layers, L, B, T, 6, 6, B * 6, SCP_F32, d_w, 0, 1e-5f, NULL, d_y, SCP_F32, st) == SCP_ERR_UNSUPPORTED, "D not a 16-byte multiple");
  EXPECT(scp_wsum_fwd(d_layers, L, B, T, D, D, B * D, SCP_F32, d_w, SCP_NORM_UTT_MEAN, 1e-5f, NULL, d_y, SCP_F32, st) == SCP_ERR_INVALID, "method2 without utt_scale");
  EXPECT(scp_wsum_bwd(d_layers, L, B, T, D, D, B * D, SCP_F32, d_w, 0, 1e-5f, NULL, d_gy, SCP_F32, d_dw, NULL, d_ws, ws_bytes ? ws_bytes - 1 : 0, st) ==
             (ws_bytes ? SCP_ERR_WORKSPACE : SCP_OK), "short workspace");
  EXPECT(strlen(scp_last_error_string(SCP_ERR_WORKSPACE)) > 0, "error string");
  CUDA_OK(cudaStreamSynchronize(st));
  for (int l = 0; l < L; ++l) { free(h_layers[l]); cudaFree((void*)d_layers[l]); }
  cudaFree(d_w); cudaFree(d_gy); cudaFree(d_y); cudaFree(d_dw); cudaFree(d_ws);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* S2: idx = first arg-max over the unmasked columns of cos(kw, E); keywords = E[idx]                              */
static void check_vq(cudaStream_t st) {
  enum { M = 12, K = 4, V = 700, D = 64 };
  static float table[V * D], kw[M * D];
  for (int i = 0; i < V * D; ++i) table[i] = 0.02f * gauss() + 0.003f;
  for (int i = 0; i < M * D; ++i) kw[i] = 0.02f * gauss() + 0.003f;
  memcpy(kw + 5 * D, table + 2 * D, D * 4);   /* row 5's best raw match is the masked column 2 */
  for (int d = 0; d < D; ++d) kw[7 * D + d] = 3.0f * table[(V - 1) * D + d]; /* row 7 -> last column */
  const int64_t Vp = scp_vq_padded_vocab(V), Mp = (M + 127) / 128 * 128;
  EXPECT(Vp >= V && Vp % 256 == 0, "padded vocab %lld", (long long)Vp);
  float* d_table = (float*)to_dev(table, sizeof table);
  float* d_kw = (float*)to_dev(kw, sizeof kw);
  void* d_hat = dev_alloc((size_t)Vp * D * 2);
  void* d_hat_t = dev_alloc((size_t)Vp * D * 2);
  float* d_norm = (float*)dev_alloc((size_t)Vp * 4);
  float* d_mean = (float*)dev_alloc((D + 1) * 4);
  const float tau = 0.1f;
  float* d_tau = (float*)to_dev(&tau, 4);
  int64_t* d_idx = (int64_t*)dev_alloc(M * 8);
  float* d_out = (float*)dev_alloc(M * D * 4);
  float* d_stats = (float*)dev_alloc(M * 4 * 4);
  float* d_hist = (float*)dev_alloc((size_t)Vp * 4);
  float* d_avg = (float*)dev_alloc((size_t)Vp * 4);
  float* d_metrics = (float*)dev_alloc((3 + K) * 4);
  void* d_kw_hat = dev_alloc((size_t)Mp * D * 2);
  size_t ws_bytes = scp_vq_fwd_workspace_bytes(M, V, D);
  void* d_ws = dev_alloc(ws_bytes);
  const int32_t masked[3] = {0, 2, 3};   /* prob_msk, my_vector_quantizer.py:64 */

  int rc = scp_vq_prepare_table(d_table, V, D, d_hat, d_hat_t, d_norm, d_mean, st);
  EXPECT(rc == SCP_OK, "scp_vq_prepare_table -> %d (%s)", rc, scp_last_error_string(rc));
  rc = scp_vq_fwd(d_kw, M, K, V, D, d_hat, d_norm, d_table, masked, 3, d_tau, d_idx, d_out, d_stats, d_hist, d_avg, d_metrics,
                  d_kw_hat, d_ws, ws_bytes, st);
  EXPECT(rc == SCP_OK, "scp_vq_fwd -> %d (%s)", rc, scp_last_error_string(rc));
  CUDA_OK(cudaStreamSynchronize(st));
  int64_t idx[M];
  static float out[M * D], avg[V], hist[V];
  float metrics[3 + K];
  to_host(idx, d_idx, sizeof idx);
  to_host(out, d_out, sizeof out);
  to_host(avg, d_avg, sizeof avg);
  to_host(hist, d_hist, sizeof hist);
  to_host(metrics, d_metrics, sizeof metrics);

  static double avg_ref[V];
  double hist_ref[V];
  memset(avg_ref, 0, sizeof avg_ref);
  memset(hist_ref, 0, sizeof hist_ref);
  for (int m = 0; m < M; ++m) {
    double nk = 0, best = -2, z = 0;
    static double c[V];
    int arg = -1;
    for (int d = 0; d < D; ++d) nk += (double)kw[m * D + d] * kw[m * D + d];
    nk = fmax(sqrt(nk), 1e-8);
    for (int v = 0; v < V; ++v) {
      double dot = 0, nv = 0;
      for (int d = 0; d < D; ++d) { dot += (double)kw[m * D + d] * table[v * D + d]; nv += (double)table[v * D + d] * table[v * D + d]; }
      c[v] = dot / (nk * fmax(sqrt(nv), 1e-8));
      if (v == 0 || v == 2 || v == 3) continue;
      z += exp(c[v]);
      if (c[v] > best) { best = c[v]; arg = v; }
    }
    for (int v = 0; v < V; ++v)
      if (!(v == 0 || v == 2 || v == 3)) avg_ref[v] += exp(c[v]) / z / M;
    hist_ref[arg] += 1;
    EXPECT(idx[m] == arg, "vq row %d: idx %lld, exact arg-max %d", m, (long long)idx[m], arg);
    EXPECT(memcmp(out + m * D, table + (size_t)idx[m] * D, D * 4) == 0, "vq row %d: keywords != E[idx]", m);
  }
  EXPECT(idx[5] != 2 && idx[7] == V - 1, "masked best match / last column: %lld %lld", (long long)idx[5], (long long)idx[7]);
  double code_ppl = 0, prob_ppl = 0;
  for (int v = 0; v < V; ++v) {
    EXPECT(fabs(avg[v] - avg_ref[v]) < 1e-3 * (1.0 / V), "avg_probs[%d] %.6g vs %.6g", v, avg[v], avg_ref[v]);
    EXPECT(hist[v] == (float)hist_ref[v], "code_hist[%d]", v);
    code_ppl -= hist_ref[v] / M * log(hist_ref[v] / M + 1e-7);   /* my_vector_quantizer.py:94-99 */
    prob_ppl -= avg_ref[v] * log(avg_ref[v] + 1e-7);              /* :119-121 */
  }
  EXPECT(fabs(metrics[0] - exp(code_ppl)) < 1e-3 * exp(code_ppl), "code_perplexity %.5f vs %.5f", metrics[0], exp(code_ppl));
  EXPECT(fabs(metrics[1] - exp(prob_ppl)) < 1e-3 * exp(prob_ppl), "prob_perplexity %.5f vs %.5f", metrics[1], exp(prob_ppl));

  EXPECT(scp_vq_fwd(d_kw, M, K, V, 48, d_hat, d_norm, d_table, masked, 3, d_tau, d_idx, d_out, d_stats, d_hist, d_avg, d_metrics, d_kw_hat, d_ws, ws_bytes, st) == SCP_ERR_UNSUPPORTED, "D = 48");
  EXPECT(scp_vq_fwd(d_kw, M, 5, V, D, d_hat, d_norm, d_table, masked, 3, d_tau, d_idx, d_out, d_stats, d_hist, d_avg, d_metrics, d_kw_hat, d_ws, ws_bytes, st) == SCP_ERR_INVALID, "M not a multiple of K");
  EXPECT(scp_vq_fwd(d_kw, M, K, V, D, d_hat, d_norm, d_table, masked, 3, d_tau, NULL, d_out, d_stats, d_hist, d_avg, d_metrics, d_kw_hat, d_ws, ws_bytes, st) == SCP_ERR_INVALID, "null idx");
  EXPECT(scp_vq_fwd(d_kw, M, K, V, D, d_hat, d_norm, d_table, masked, SCP_MAX_MASKED + 1, d_tau, d_idx, d_out, d_stats, d_hist, d_avg, d_metrics, d_kw_hat, d_ws, ws_bytes, st) == SCP_ERR_INVALID, "too many masked columns");
  EXPECT(scp_vq_fwd(d_kw, M, K, V, D, d_hat, d_norm, d_table, masked, 3, d_tau, d_idx, d_out, d_stats, d_hist, d_avg, d_metrics, d_kw_hat, d_ws, 16, st) == SCP_ERR_WORKSPACE, "short workspace");
  CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(d_table); cudaFree(d_kw); cudaFree(d_hat); cudaFree(d_hat_t); cudaFree(d_norm); cudaFree(d_mean); cudaFree(d_tau);
  cudaFree(d_idx); cudaFree(d_out); cudaFree(d_stats); cudaFree(d_hist); cudaFree(d_avg); cudaFree(d_metrics); cudaFree(d_kw_hat);
  cudaFree(d_ws);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* S3: masked two-way InfoNCE (losses.py:185-245), forward value, both LSE vectors and dA of the local rows        */
static void check_nce(cudaStream_t st) {
  enum { N = 24, D = 64 };
  static float A[N * D], Bm[N * D];
  int64_t ids[N];
  for (int i = 0; i < N; ++i) {
    double na = 0, nb = 0;
    for (int d = 0; d < D; ++d) { Bm[i * D + d] = gauss(); A[i * D + d] = gauss() + Bm[i * D + d]; }
    for (int d = 0; d < D; ++d) { na += (double)A[i * D + d] * A[i * D + d]; nb += (double)Bm[i * D + d] * Bm[i * D + d]; }
    for (int d = 0; d < D; ++d) { A[i * D + d] /= (float)sqrt(na); Bm[i * D + d] /= (float)sqrt(nb); }
    ids[i] = i % 9;   /* same-image pairs are masked out of the negatives (:207-214) */
  }
  const float log_scale = (float)log(1.0 / 0.07), one = 1.0f;
  float* d_A = (float*)to_dev(A, sizeof A);
  float* d_B = (float*)to_dev(Bm, sizeof Bm);
  int64_t* d_ids = (int64_t*)to_dev(ids, sizeof ids);
  float* d_ls = (float*)to_dev(&log_scale, 4);
  float* d_one = (float*)to_dev(&one, 4);
  float* d_loss = (float*)dev_alloc(4);
  float* d_lr = (float*)dev_alloc(N * 4);
  float* d_lc = (float*)dev_alloc(N * 4);
  float* d_dA = (float*)dev_alloc(N * D * 4);
  float* d_dls = (float*)dev_alloc(4);
  size_t ws_bytes = scp_nce_workspace_bytes(N, D);
  void* d_ws = dev_alloc(ws_bytes);
  int rc = scp_nce_fwd(d_A, d_B, d_ids, N, D, d_ls, 0.f, 0.f, 0, 1, 1, 1, d_loss, d_lr, d_lc, d_ws, ws_bytes, st);
  EXPECT(rc == SCP_OK, "scp_nce_fwd -> %d (%s)", rc, scp_last_error_string(rc));
  rc = scp_nce_bwd(d_A, d_B, d_ids, N, D, d_ls, 0.f, 0.f, 0, 1, 1, d_lr, d_lc, d_one, 0, N, 1, d_dA, NULL, d_dls, d_ws, ws_bytes, st);
  EXPECT(rc == SCP_OK, "scp_nce_bwd -> %d (%s)", rc, scp_last_error_string(rc));
  CUDA_OK(cudaStreamSynchronize(st));
  float loss, lr[N], lc[N], dls;
  static float dA[N * D];
  to_host(&loss, d_loss, 4); to_host(lr, d_lr, sizeof lr); to_host(lc, d_lc, sizeof lc); to_host(dA, d_dA, sizeof dA);
  to_host(&dls, d_dls, 4);

  static double S[N][N], E[N][N], zr[N], zc[N];
  const double scale = exp((double)log_scale);
  memset(zr, 0, sizeof zr); memset(zc, 0, sizeof zc);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double dot = 0;
      for (int d = 0; d < D; ++d) dot += (double)A[i * D + d] * Bm[j * D + d];
      S[i][j] = dot * scale;
      E[i][j] = (ids[i] != ids[j] || i == j) ? exp(S[i][j]) : 0.0;
      zr[i] += E[i][j]; zc[j] += E[i][j];
    }
  double ref = 0, dls_ref = 0, err = 0, nrm = 0;
  for (int i = 0; i < N; ++i) {
    ref += 0.5 * ((-S[i][i] + log(zr[i])) + (-S[i][i] + log(zc[i]))) / N;
    EXPECT(fabs(lr[i] - log(zr[i])) < 1e-4 && fabs(lc[i] - log(zc[i])) < 1e-4, "lse[%d]: %.5f %.5f vs %.5f %.5f", i, lr[i], lc[i], log(zr[i]), log(zc[i]));
  }
  /* north_star tolerance 1e-3; the loss here is a small difference of O(10) logits (strong positives), observed 2e-4 */
  EXPECT(fabs(loss - ref) < 1e-3 * fabs(ref), "nce loss %.6f vs %.6f", loss, ref);
  for (int i = 0; i < N; ++i)
    for (int d = 0; d < D; ++d) {
      double g = 0;
      for (int j = 0; j < N; ++j) {
        const double G = (E[i][j] / zr[i] + E[i][j] / zc[j]) / (2.0 * N) - (i == j ? 1.0 / N : 0.0);
        g += scale * G * Bm[j * D + d];
        if (d == 0) dls_ref += G * S[i][j];
      }
      err += (g - dA[i * D + d]) * (g - dA[i * D + d]);
      nrm += g * g;
    }
  EXPECT(sqrt(err / nrm) < 1e-3, "nce dA: relative l2 error %.3g", sqrt(err / nrm));
  EXPECT(fabs(dls - dls_ref) < 1e-3 * (fabs(dls_ref) + 1e-3), "nce d_log_scale %.6f vs %.6f", dls, dls_ref);

  EXPECT(scp_nce_fwd(d_A, d_B, d_ids, N, 48, d_ls, 0.f, 0.f, 0, 1, 1, 1, d_loss, d_lr, d_lc, d_ws, ws_bytes, st) == SCP_ERR_UNSUPPORTED, "D = 48");
  EXPECT(scp_nce_fwd(d_A, d_B, d_ids, N, D, d_ls, 0.f, 0.f, 0, 0, 0, 1, d_loss, d_lr, d_lc, d_ws, ws_bytes, st) == SCP_ERR_INVALID, "a2b = b2a = 0");
  EXPECT(scp_nce_fwd(d_A, d_B, d_ids, N, D, d_ls, 0.f, 0.f, 0, 1, 1, 1, d_loss, d_lr, d_lc, d_ws, 8, st) == SCP_ERR_WORKSPACE, "short workspace");
  EXPECT(scp_nce_bwd(d_A, d_B, d_ids, N, D, d_ls, 0.f, 0.f, 0, 1, 1, d_lr, d_lc, d_one, 5, 5, 0, d_dA, NULL, NULL, d_ws, ws_bytes, st) == SCP_ERR_INVALID, "empty row range");
  CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(d_A); cudaFree(d_B); cudaFree(d_ids); cudaFree(d_ls); cudaFree(d_one); cudaFree(d_loss); cudaFree(d_lr); cudaFree(d_lc);
  cudaFree(d_dA); cudaFree(d_dls); cudaFree(d_ws);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* S2 with saved soft-max numerators: scp_vq_fwd_save + scp_vq_bwd_saved must reproduce scp_vq_fwd + scp_vq_bwd          */
static void check_vq_saved(cudaStream_t st) {
  enum { M = 256, K = 8, V = 700, D = 64 };   /* two 128-row tiles: the saved path is available */
  static float table[V * D], kw[M * D], g[M * D];
  for (int i = 0; i < V * D; ++i) table[i] = 0.02f * gauss() + 0.003f;
  for (int i = 0; i < M * D; ++i) { kw[i] = 0.02f * gauss() + 0.003f; g[i] = gauss(); }
  for (int d = 0; d < D; ++d) kw[9 * D + d] = 2.0f * table[17 * D + d];   /* a row whose cosine with column 17 is 1 */
  const int64_t Vp = scp_vq_padded_vocab(V), Mp = (M + 127) / 128 * 128;
  EXPECT(scp_vq_bwd_saved_available(M, V, D) == 1 && scp_vq_bwd_saved_available(12, V, D) == 0 &&
             scp_vq_bwd_saved_available(M, V, 768) == 1, "scp_vq_bwd_saved_available");
  EXPECT(scp_vq_saved_probs_bytes(M, V) == (size_t)Mp * Vp * 2, "scp_vq_saved_probs_bytes");
  EXPECT(scp_vq_fwd_save_workspace_bytes(M, V, D) + (size_t)Mp * Vp * 2 <= scp_vq_fwd_workspace_bytes(M, V, D) + 4096,
         "the saved forward needs no (M,V) scratch inside its workspace");
  float* d_table = (float*)to_dev(table, sizeof table);
  float* d_kw = (float*)to_dev(kw, sizeof kw);
  float* d_g = (float*)to_dev(g, sizeof g);
  void* d_hat = dev_alloc((size_t)Vp * D * 2);
  void* d_hat_t = dev_alloc((size_t)Vp * D * 2);
  float* d_norm = (float*)dev_alloc((size_t)Vp * 4);
  float* d_mean = (float*)dev_alloc((D + 1) * 4);
  const float tau = 0.1f;
  float* d_tau = (float*)to_dev(&tau, 4);
  const int32_t masked[3] = {0, 2, 3};
  int rc = scp_vq_prepare_table(d_table, V, D, d_hat, d_hat_t, d_norm, d_mean, st);
  EXPECT(rc == SCP_OK, "scp_vq_prepare_table -> %d", rc);
  static int64_t idx[2][M];
  static float avg[2][V], gk[2][M * D];
  for (int saved = 0; saved < 2; ++saved) {
    int64_t* d_idx = (int64_t*)dev_alloc(M * 8);
    float* d_out = (float*)dev_alloc(M * D * 4);
    float* d_stats = (float*)dev_alloc(M * 4 * 4);
    float* d_hist = (float*)dev_alloc((size_t)Vp * 4);
    float* d_avg = (float*)dev_alloc((size_t)Vp * 4);
    float* d_metrics = (float*)dev_alloc((3 + K) * 4);
    void* d_kw_hat = dev_alloc((size_t)Mp * D * 2);
    float* d_gk = (float*)dev_alloc(M * D * 4);
    void* d_saved = saved ? dev_alloc(scp_vq_saved_probs_bytes(M, V)) : NULL;
    const size_t fws = saved ? scp_vq_fwd_save_workspace_bytes(M, V, D) : scp_vq_fwd_workspace_bytes(M, V, D);
    const size_t bws = saved ? scp_vq_bwd_saved_workspace_bytes(M, V, D, 0) : scp_vq_bwd_workspace_bytes(M, V, D);
    void* d_fws = dev_alloc(fws);
    void* d_bws = dev_alloc(bws);
    rc = scp_vq_fwd_save(d_kw, M, K, V, D, d_hat, d_norm, d_table, masked, 3, d_tau, d_idx, d_out, d_stats, d_hist, d_avg,
                         d_metrics, d_kw_hat, d_saved, d_fws, fws, st);
    EXPECT(rc == SCP_OK, "scp_vq_fwd_save(saved=%d) -> %d (%s)", saved, rc, scp_last_error_string(rc));
    rc = scp_vq_bwd_saved(d_g, d_kw, M, V, D, d_kw_hat, d_hat, d_hat_t, d_norm, d_mean, d_stats, masked, 3, d_tau, d_saved,
                          d_gk, NULL, d_bws, bws, st);
    EXPECT(rc == SCP_OK, "scp_vq_bwd_saved(saved=%d) -> %d (%s)", saved, rc, scp_last_error_string(rc));
    CUDA_OK(cudaStreamSynchronize(st));
    to_host(idx[saved], d_idx, sizeof idx[saved]);
    to_host(avg[saved], d_avg, sizeof avg[saved]);
    to_host(gk[saved], d_gk, sizeof gk[saved]);
    cudaFree(d_idx); cudaFree(d_out); cudaFree(d_stats); cudaFree(d_hist); cudaFree(d_avg); cudaFree(d_metrics);
    cudaFree(d_kw_hat); cudaFree(d_gk); cudaFree(d_fws); cudaFree(d_bws);
    if (d_saved) cudaFree(d_saved);
  }
  EXPECT(memcmp(idx[0], idx[1], sizeof idx[0]) == 0 && idx[1][9] == 17, "saved forward: same arg-max codes");
  double num = 0, den = 0;
  for (int v = 0; v < V; ++v) EXPECT(fabs(avg[0][v] - avg[1][v]) < 1e-3 * (1.0 / V), "avg_probs[%d] %.6g vs %.6g", v, avg[1][v], avg[0][v]);
  for (int i = 0; i < M * D; ++i) { num += ((double)gk[1][i] - gk[0][i]) * ((double)gk[1][i] - gk[0][i]); den += (double)gk[0][i] * gk[0][i]; }
  EXPECT(den > 0 && sqrt(num / den) < 1e-3, "keyword gradient: saved vs recompute, relative error %.3e", sqrt(num / fmax(den, 1e-300)));
  cudaFree(d_table); cudaFree(d_kw); cudaFree(d_g); cudaFree(d_hat); cudaFree(d_hat_t); cudaFree(d_norm); cudaFree(d_mean); cudaFree(d_tau);
}

int main(void) {
  int n_dev = 0;
  printf("libscp_b200 version %d\n", scp_version());
  if (scp_version() <= 0) return 1;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    fprintf(stderr, "no CUDA device: nothing to run (the library has no CPU path)\n");
    return 77;
  }
  CUDA_OK(cudaSetDevice(0));
  cudaStream_t st;
  CUDA_OK(cudaStreamCreate(&st));
  const int before = scp_num_launches();
  check_wsum(st);
  check_vq(st);
  check_vq_saved(st);
  check_nce(st);
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaGetLastError());
  printf("%d kernels launched, %d failure(s)\n", scp_num_launches() - before, failures);
  return failures ? 1 : 0;
}
