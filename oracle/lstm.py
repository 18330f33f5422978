"""Stage 2 of the path: the MC-dropout LSTM regressor (``estimate/nn_models.py:160-207``) on the CPU.

The arithmetic lives in torch (``torch.nn.LSTM(batch_first=True, dropout=p)`` + ``torch.nn.Linear``,
nn_models.py:169-174), not under /root/reference, so it is restated here from the published cell equations

    g   = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh          (gate rows ordered i, f, g~, o)
    i,f,o = sigmoid(.), g~ = tanh(.) ; c_t = f*c_{t-1} + i*g~ ; h_t = o*tanh(c_t) ; h_0 = c_0 = 0

with, in train mode, every non-final layer's output sequence multiplied by a Bernoulli(1-p) mask scaled by
1/(1-p).  ``forward_with_masks`` is the decomposed float32 restatement (parity anchor,
checked against ``torch.nn.LSTM`` in tests/golden/make_golden.py and tests/test_oracle_golden.py);
``TorchDropoutLSTM`` is the same module the reference builds, used for mask replay and as the timed CPU
baseline.  Test infrastructure only.
"""
import numpy as np
import torch


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x, dtype=np.float32))).astype(np.float32)


def lstm_dims(state):
    """(I, H, L, O) from reference state-dict key shapes (``lstm.weight_ih_l{k}``, ``output_layer.weight``)."""
    H4, I = state["lstm.weight_ih_l0"].shape
    L = sum(1 for k in state if k.startswith("lstm.weight_ih_l"))
    return int(I), int(H4 // 4), int(L), int(state["output_layer.weight"].shape[0])


def forward_with_masks(state, x, masks=None, p=0.2, last_only=False, return_hidden=False):
    """Decomposed forward pass.  ``x (rows, T, I)`` float32; ``masks``: list of L-1 arrays ``(T, rows, H)`` of
    {0,1} (time-major, the order torch draws them) or ``None`` for eval mode.  Returns ``(rows, T, O)`` like
    ``DropoutLSTM.forward`` (nn_models.py:180-189), or ``(rows, O)`` of the last step if ``last_only``."""
    st = {k: np.asarray(v, dtype=np.float32) for k, v in state.items()}
    I, H, L, O = lstm_dims(st)
    seq = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    rows, T, _ = seq.shape
    for l in range(L):
        w_ih, w_hh = st[f"lstm.weight_ih_l{l}"], st[f"lstm.weight_hh_l{l}"]
        bias = st[f"lstm.bias_ih_l{l}"] + st[f"lstm.bias_hh_l{l}"]
        h = np.zeros((rows, H), np.float32)
        c = np.zeros((rows, H), np.float32)
        out = np.empty((rows, T, H), np.float32)
        for t in range(T):
            g = seq[:, t, :] @ w_ih.T + h @ w_hh.T + bias
            gi, gf = _sigmoid(g[:, :H]), _sigmoid(g[:, H:2 * H])
            gg, go = np.tanh(g[:, 2 * H:3 * H]), _sigmoid(g[:, 3 * H:])
            c = gf * c + gi * gg
            h = go * np.tanh(c)
            out[:, t, :] = h
        if l < L - 1 and masks is not None:
            m = np.asarray(masks[l], dtype=np.float32).transpose(1, 0, 2)      # (rows, T, H)
            out = out * m / np.float32(1.0 - p)
        seq = out
    if return_hidden:
        return seq
    w_o, b_o = st["output_layer.weight"], st["output_layer.bias"]
    if last_only:
        return seq[:, -1, :] @ w_o.T + b_o
    return seq @ w_o.T + b_o


def replay_torch_masks(seed, T, rows, H, L, p):
    """Masks ``torch.nn.LSTM`` (CPU, train mode) draws after ``torch.manual_seed(seed)``: for gap
    l = 0..L-2 one ``torch.empty(T, rows, H).bernoulli_(1-p)`` and nothing else."""
    torch.manual_seed(seed)
    return [torch.empty(T, rows, H).bernoulli_(1 - p).numpy().astype(np.uint8) for _ in range(L - 1)]


class TorchDropoutLSTM(torch.nn.Module):
    """Same construction as ``DropoutLSTM`` (nn_models.py:160-178) so reference state dicts load unchanged."""

    def __init__(self, input_size, hidden_layer_size, hidden_layer_count, output_size, dropout=0.2):
        super().__init__()
        self.lstm = torch.nn.LSTM(input_size, hidden_size=hidden_layer_size, num_layers=hidden_layer_count,
                                  batch_first=True, dropout=dropout)
        self.output_layer = torch.nn.Linear(hidden_layer_size, output_size)

    def forward(self, x, hs=None):
        seq, _ = self.lstm(x, hs)
        return self.output_layer(seq)

    def monte_carlo_predictions(self, n_samples, x, hs=None):
        # nn_models.py:191-207: batch>1 refused, only the LSTM sub-module goes to train mode, input tiled n times
        if x.shape[0] > 1:
            raise UserWarning("MC predictions only for batch size 1")
        self.lstm.train()
        return self(x.repeat((n_samples, 1, 1)), hs)

    @classmethod
    def from_state(cls, state, p=0.2):
        I, H, L, O = lstm_dims(state)
        m = cls(I, H, L, O, p)
        m.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in state.items()})
        m.eval()
        return m
