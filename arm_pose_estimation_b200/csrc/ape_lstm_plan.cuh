// Shared between the fp32 (ape_lstm_fma.cu) and tensor-core (ape_lstm_tc.cu) MC-LSTM paths: the tiling plan of the
// fp32 layer kernels and the launcher of one fp32 layer (the tensor-core path runs layer 0 through it).
#pragma once
#include "ape_common.cuh"

namespace ape {

struct FmaPlan {
    int rt0, rt1;                     // rows per tile / 16 of layer 0 and of layers >= 1
    long long tiles0, tiles1;
    size_t seq0_bytes, seq1_bytes, total;
};

int check_lstm_args(const ape_lstm_args* g);
int make_plan(int I, int H, int L, int T, long long E, int n, FmaPlan* p);
int fma_launch_layer(const ape_lstm_args* g, int l, const FmaPlan& p, const float* seq_in, float* seq_out, cudaStream_t st);

}  // namespace ape
