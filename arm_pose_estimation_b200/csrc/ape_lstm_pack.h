// Layout of the packed fp32 weight blob consumed by the MC-LSTM kernels (plain C, host + device).
//
// The reference keeps torch.nn.LSTM / Linear parameters under the state-dict keys lstm.weight_ih_l{k} (4H, I|H),
// lstm.weight_hh_l{k} (4H, H), lstm.bias_ih_l{k}, lstm.bias_hh_l{k} (4H), output_layer.weight (O, H),
// output_layer.bias (O) with gate rows ordered i, f, g, o (nn_models.py:169-174).  The kernels want, per layer,
// ONE K-major matrix  Wp[k][c]  (K = Kin_pad + H rows: first the input weights, zero-padded to a multiple of
// APE_KSLICE rows, then the recurrent weights) whose column c = 4*u + g interleaves the four gates of hidden
// unit u, followed by the summed bias bp[c] = b_ih[g*H+u] + b_hh[g*H+u].  After the layers: W_o (O, H) row-major
// and b_o (O).  arm_pose_estimation_b200.estimate.nn_models.pack_lstm_weights() writes exactly this layout.
#ifndef APE_LSTM_PACK_H
#define APE_LSTM_PACK_H

#include <stdint.h>

#define APE_KSLICE 16

#ifdef __CUDACC__
#define APE_PACK_HD __host__ __device__ __forceinline__
#else
#define APE_PACK_HD static inline
#endif

APE_PACK_HD int ape_pack_kin_pad(int layer, int I, int H) {
    return layer == 0 ? ((I + APE_KSLICE - 1) / APE_KSLICE) * APE_KSLICE : H;
}
// float offset of layer `layer`'s Wp inside the blob
APE_PACK_HD int64_t ape_pack_layer_offset(int layer, int I, int H) {
    int64_t off = 0;
    for (int l = 0; l < layer; ++l) off += (int64_t)(ape_pack_kin_pad(l, I, H) + H) * 4 * H + 4 * H;
    return off;
}
APE_PACK_HD int64_t ape_pack_bias_offset(int layer, int I, int H) {
    return ape_pack_layer_offset(layer, I, H) + (int64_t)(ape_pack_kin_pad(layer, I, H) + H) * 4 * H;
}
APE_PACK_HD int64_t ape_pack_out_offset(int I, int H, int L) { return ape_pack_layer_offset(L, I, H); }
APE_PACK_HD int64_t ape_pack_total_floats(int I, int H, int L, int O) {
    return ape_pack_out_offset(I, H, L) + (int64_t)O * H + O;
}

#endif  // APE_LSTM_PACK_H
