/*
 * A plain-C consumer of the C ABI (include/ape_b200.h): no Python, no torch - what a C / C++ / cgo / JNI host would do.
 * One call of the three stages for B streams x nF frames x n MC samples of the watch-only model shape
 * (stage 1 ape_features -> stage 2 ape_mc_lstm_fma, the exact fp32 kernel -> stage 3 ape_fk_reduce), on inputs it generates
 * itself, and a dump of inputs + outputs so that tests/test_cabi_consumer.py can replay the same call through the ctypes binding and
 * compare bit for bit.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/cabi_consumer.c -o build/cabi_consumer \
 *       -Larm_pose_estimation_b200/lib -lape_b200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/arm_pose_estimation_b200/lib
 *   build/cabi_consumer out.bin
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ape_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define APE(x) do { int r_ = (x); if (r_ != APE_OK) { fprintf(stderr, "%s -> %d (%s)\n", #x, r_, ape_last_cuda_error()); return 3; } } while (0)

static uint32_t lcg_state = 12345u;
static float lcg_uniform(void) {                       /* [0, 1) */
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return (float)(lcg_state >> 8) * (1.0f / 16777216.0f);
}

int main(int argc, char** argv) {
    /* the watch-only model of the reference: 28-float rows, I = 20, H = 256, L = 2, T = 8, O = 12 */
    enum { B = 3, NF = 2, N = 5, NCOLS = 28, I = 20, H = 256, L = 2, T = 8, O = 12, SMOOTH = 2 };
    const int E = B * NF, S = SMOOTH * N, feat_ring = NF + T - 1, pred_ring = NF + SMOOTH - 1;
    int64_t blob_floats = 0;
    uint64_t ws_bytes = 0;
    int sm = 0, smem = 0, cc_major = 0, cc_minor = 0;
    APE(ape_device_info(&sm, &smem, &cc_major, &cc_minor));
    APE(ape_lstm_blob_floats(I, H, L, O, &blob_floats));
    APE(ape_mc_lstm_workspace_bytes(I, H, L, T, O, E, N, &ws_bytes));

    /* host inputs: raw wire rows (quaternion columns need not be unit: the kernels normalise like the reference), weights, stats */
    float* raw = (float*)malloc(sizeof(float) * E * NCOLS);
    float* blob = (float*)malloc(sizeof(float) * blob_floats);
    double xx_m[I], xx_s[I];
    float yy_m[O], yy_s[O], body9[9] = {-0.22f, 0.0f, 0.0f, -0.26f, 0.0f, 0.0f, -0.17f, 0.43f, -0.01f};
    for (int i = 0; i < E * NCOLS; ++i) raw[i] = 2.0f * lcg_uniform() - 1.0f;
    for (int64_t i = 0; i < blob_floats; ++i) blob[i] = 0.12f * (lcg_uniform() - 0.5f);
    for (int i = 0; i < I; ++i) { xx_m[i] = 0.1 * i - 1.0; xx_s[i] = 0.5 + 0.05 * i; }
    for (int i = 0; i < O; ++i) { yy_m[i] = 0.05f * i; yy_s[i] = 0.8f + 0.02f * i; }

    float *d_raw, *d_blob, *d_feats, *d_preds, *d_yy_m, *d_yy_s, *d_body, *d_msg, *d_samples, *d_std;
    double *d_xx_m, *d_xx_s;
    int32_t* d_status;
    void* d_ws;
    CK(cudaMalloc((void**)&d_raw, sizeof(float) * E * NCOLS));
    CK(cudaMalloc((void**)&d_blob, sizeof(float) * blob_floats));
    CK(cudaMalloc((void**)&d_feats, sizeof(float) * B * feat_ring * I));
    CK(cudaMalloc((void**)&d_preds, sizeof(float) * B * pred_ring * N * O));
    CK(cudaMalloc((void**)&d_xx_m, sizeof(double) * I));
    CK(cudaMalloc((void**)&d_xx_s, sizeof(double) * I));
    CK(cudaMalloc((void**)&d_yy_m, sizeof(float) * O));
    CK(cudaMalloc((void**)&d_yy_s, sizeof(float) * O));
    CK(cudaMalloc((void**)&d_body, sizeof(float) * 9));
    CK(cudaMalloc((void**)&d_msg, sizeof(float) * E * 25));
    CK(cudaMalloc((void**)&d_samples, sizeof(float) * E * S * 6));
    CK(cudaMalloc((void**)&d_std, sizeof(float) * E * 6));
    CK(cudaMalloc((void**)&d_status, sizeof(int32_t) * E));
    CK(cudaMalloc(&d_ws, ws_bytes + 256));
    CK(cudaMemset(d_feats, 0, sizeof(float) * B * feat_ring * I));
    CK(cudaMemset(d_preds, 0, sizeof(float) * B * pred_ring * N * O));
    CK(cudaMemcpy(d_raw, raw, sizeof(float) * E * NCOLS, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_blob, blob, sizeof(float) * blob_floats, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_xx_m, xx_m, sizeof(xx_m), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_xx_s, xx_s, sizeof(xx_s), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_yy_m, yy_m, sizeof(yy_m), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_yy_s, yy_s, sizeof(yy_s), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_body, body9, sizeof(body9), cudaMemcpyHostToDevice));

    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    /* stage 1: rows of frames 0 .. NF-1 of every stream -> feature ring */
    APE(ape_features(d_raw, APE_LAYOUT_WATCH_ONLY, APE_KIND_WATCH_ONLY, d_xx_m, d_xx_s, 1, d_feats, B, NF, 0, NULL, feat_ring, st));
    /* stage 2: MC-dropout LSTM with the library's Philox masks -> prediction ring */
    ape_lstm_args a;
    memset(&a, 0, sizeof(a));
    a.weights = d_blob; a.I = I; a.H = H; a.L = L; a.T = T; a.O = O; a.dropout_p = 0.2f;
    a.feat_ring_buf = d_feats; a.feat_ring = feat_ring; a.B = B; a.nF = NF; a.frame0 = 0; a.n_samples = N;
    a.mask_mode = APE_MASK_PHILOX; a.philox_seed = 0x1234abcdULL; a.stream_id0 = 7;
    a.workspace = (void*)(((uintptr_t)d_ws + 255) & ~(uintptr_t)255);
    a.preds = d_preds; a.pred_ring = pred_ring;
    APE(ape_mc_lstm_fma(&a, st));
    /* stage 3: de-normalise, 6D -> quaternion, forward kinematics, MC mean / std over the smoothing window */
    APE(ape_fk_reduce(d_preds, pred_ring, d_yy_m, d_yy_s, d_body, APE_TARGET_ORI_CAL_LARM_UARM, O, B, NF, 0, NULL, N, SMOOTH,
                      d_msg, d_samples, d_std, NULL, d_status, st));
    CK(cudaStreamSynchronize(st));

    float* msg = (float*)malloc(sizeof(float) * E * 25);
    float* std6 = (float*)malloc(sizeof(float) * E * 6);
    float* samples = (float*)malloc(sizeof(float) * E * S * 6);
    int32_t status[E];
    CK(cudaMemcpy(msg, d_msg, sizeof(float) * E * 25, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(std6, d_std, sizeof(float) * E * 6, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(samples, d_samples, sizeof(float) * E * S * 6, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(status, d_status, sizeof(status), cudaMemcpyDeviceToHost));

    double qn_err = 0.0;
    for (int e = 0; e < E; ++e) {
        const float* q = msg + e * 25;
        qn_err = fmax(qn_err, fabs(sqrt((double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2] + (double)q[3] * q[3]) - 1.0));
        if (status[e] != 0) { fprintf(stderr, "estimate %d: degenerate 6D pair\n", e); return 4; }
    }
    printf("abi %d, sm_%d%d, %d SMs; %d estimates x %d rows: hand of estimate 0 = (%.6f, %.6f, %.6f), max | |q| - 1 | = %.2e\n",
           ape_abi_version(), cc_major, cc_minor, sm, E, S, msg[4], msg[5], msg[6], qn_err);
    if (!(qn_err < 1e-5)) return 5;

    if (argc > 1) {     /* dump: header of int32 sizes, then raw, blob, xx_m, xx_s, yy_m, yy_s, body9, msg, std, samples */
        FILE* f = fopen(argv[1], "wb");
        if (!f) return 6;
        const int32_t hdr[12] = {B, NF, N, NCOLS, I, H, L, T, O, SMOOTH, (int32_t)blob_floats, 0};
        fwrite(hdr, sizeof(hdr), 1, f);
        fwrite(raw, sizeof(float), (size_t)E * NCOLS, f);
        fwrite(blob, sizeof(float), (size_t)blob_floats, f);
        fwrite(xx_m, sizeof(double), I, f);
        fwrite(xx_s, sizeof(double), I, f);
        fwrite(yy_m, sizeof(float), O, f);
        fwrite(yy_s, sizeof(float), O, f);
        fwrite(body9, sizeof(float), 9, f);
        fwrite(msg, sizeof(float), (size_t)E * 25, f);
        fwrite(std6, sizeof(float), (size_t)E * 6, f);
        fwrite(samples, sizeof(float), (size_t)E * S * 6, f);
        fclose(f);
    }
    return 0;
}
