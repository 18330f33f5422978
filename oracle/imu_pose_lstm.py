"""``ImuPoseLSTM`` on the CPU (``estimate/nn_models.py:210-249``): ``Linear(input, 256) + relu`` in front of a plain 2-layer
LSTM(256, 256) and ``Linear(256, output)`` on every step.  numpy float32 restatement on top of ``oracle/lstm.py``; the
reference's ``monte_carlo_predictions`` (``:247-252``) is a regular forward pass.  Test infrastructure only.
"""
import numpy as np

from oracle import lstm as OL


def forward(state, x):
    """``x (rows, T, I)`` -> ``(rows, T, O)`` in eval mode (no inter-layer dropout)."""
    st = {k: np.asarray(v, dtype=np.float32) for k, v in state.items()}
    a = np.asarray(x, dtype=np.float32) @ st["input_layer.weight"].T + st["input_layer.bias"]      # nn_models.py:243
    a = np.maximum(a, np.float32(0.0)).astype(np.float32)                                          # F.relu
    core = {k: v for k, v in st.items() if not k.startswith("input_layer.")}
    return OL.forward_with_masks(core, a)                                                           # :244-245
