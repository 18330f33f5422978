"""The MC-dropout feed-forward regressors on the CPU (``estimate/nn_models.py:252-370``): ``DropoutFF`` and ``DropoutFF2D``.

Linear + ``leaky_relu`` (slope 0.01, torch's default) input and hidden layers, ``Dropout(p)`` in front of the output layer only
(nn_models.py:351-353 / :290-292); ``monte_carlo_predictions`` repeats the input n times with the dropout active.  numpy
float32 restatement with injected masks ``(rows, H)`` of {0,1}.  Test infrastructure only.
"""
import numpy as np


def ff_dims(state):
    """(I, H, Lh, O) from the reference's state-dict keys (``_input_layer``, ``_hidden_layers.{k}``, ``_output_layer``)."""
    H, I = state["_input_layer.weight"].shape
    Lh = sum(1 for k in state if k.startswith("_hidden_layers.") and k.endswith(".weight"))
    return int(I), int(H), int(Lh), int(state["_output_layer.weight"].shape[0])


def hidden_stack(state, x):
    st = {k: np.asarray(v, dtype=np.float32) for k, v in state.items()}
    _, _, Lh, _ = ff_dims(st)
    act = lambda v: np.where(v > 0, v, np.float32(0.01) * v).astype(np.float32)
    h = act(np.asarray(x, dtype=np.float32) @ st["_input_layer.weight"].T + st["_input_layer.bias"])
    for l in range(Lh):
        h = act(h @ st[f"_hidden_layers.{l}.weight"].T + st[f"_hidden_layers.{l}.bias"])
    return h


def forward_with_masks(state, x, masks=None, p=0.2):
    """``x (rows, I)`` (FF2D: flattened) -> ``(rows, O)``; ``masks (rows, H)`` or ``None`` for eval mode."""
    st = {k: np.asarray(v, dtype=np.float32) for k, v in state.items()}
    h = hidden_stack(st, x)
    if masks is not None:
        h = h * np.asarray(masks, dtype=np.float32) / np.float32(1.0 - p)
    return h @ st["_output_layer.weight"].T + st["_output_layer.bias"]
