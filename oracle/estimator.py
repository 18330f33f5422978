"""The per-frame estimation loop of one stream on the CPU: the three calls of ``estimator.py:174-176``.

``OracleEstimator`` restates ``Estimator`` (estimate/estimator.py:17-137) plus the ``make_prediction_from_row_hist``
of the three NN estimators (watch_only.py:84-97 and twins).  Dropout masks come either from torch's own CPU
RNG (``mask_source="torch"`` - what the reference does; replayable with ``torch.manual_seed``) or from a
callable ``mask_source(frame_idx) -> [ (T, n, H) uint8 ] * (L-1)`` (injected-mask parity runs).
Test infrastructure only.
"""
import numpy as np
import torch

from oracle import features as F
from oracle import fk
from oracle import lstm as LS


class OracleEstimator:
    def __init__(self, kind, lookup, state, stats, target, seq_len, smooth=1, mc_samples=25,
                 body=None, dropout=0.2, mask_source="torch"):
        self.kind, self.lookup, self.target = kind, lookup, target
        self.state = {k: np.asarray(v, dtype=np.float32) for k, v in state.items()}
        self.xx_m, self.xx_s = np.asarray(stats["xx_m"]), np.asarray(stats["xx_s"])
        self.yy_m, self.yy_s = np.asarray(stats["yy_m"]), np.asarray(stats["yy_s"])
        self.T, self.smooth, self.n, self.p = max(1, seq_len), max(1, smooth), mc_samples, dropout
        self.body = np.asarray(body if body is not None else
                               [[-0.22, 0, 0, -0.26, 0, 0, -0.1704612, 0.4309841, -0.00670862]], dtype=np.float64)
        self.mask_source = mask_source
        self.model = LS.TorchDropoutLSTM.from_state(self.state, dropout) if mask_source == "torch" else None
        self.reset()

    def reset(self):
        self.row_hist, self.smooth_hist, self.frame = [], [], 0

    # estimator.py:174
    def parse_row_to_xx(self, row):
        return F.parse_row(self.kind, row, self.lookup)

    # estimator.py:93-120
    def add_xx_to_row_hist_and_make_prediction(self, xx):
        self.row_hist.append(xx)
        while len(self.row_hist) < self.T:          # pad by repeating the newest row (:96-97)
            self.row_hist.append(xx)
        while len(self.row_hist) > self.T:
            del self.row_hist[0]
        xx_hist = (np.vstack(self.row_hist) - self.xx_m) / self.xx_s
        pred = self.make_prediction_from_row_hist(xx_hist) * self.yy_s + self.yy_m
        if self.smooth > 1:
            self.smooth_hist.append(pred)
            while len(self.smooth_hist) < self.smooth:
                self.smooth_hist.append(pred)
            while len(self.smooth_hist) > self.smooth:
                del self.smooth_hist[0]
            pred = np.vstack(self.smooth_hist)
        self.frame += 1
        return pred

    # watch_only.py:84-97
    def make_prediction_from_row_hist(self, xx_hist):
        if self.mask_source == "torch":
            xx = torch.tensor(xx_hist[None, :, :], dtype=torch.float32)
            with torch.no_grad():
                t_preds = self.model.monte_carlo_predictions(x=xx, n_samples=self.n)
            return t_preds.numpy()[:, -1, :]
        x = np.repeat(np.asarray(xx_hist, dtype=np.float32)[None], self.n, axis=0)
        return LS.forward_with_masks(self.state, x, self.mask_source(self.frame), self.p, last_only=True)

    # estimator.py:122-137
    def msg_from_pred(self, pred, add_mc_samples=True):
        est = fk.arm_pose_from_nn_targets(pred, self.body, self.target)
        msg = fk.msg_from_est(est, self.body, self.target)
        self.last_est = est
        if add_mc_samples:
            msg = list(msg)
            if est.shape[0] > 1:
                for e_row in est:
                    msg += list(e_row[:6])
        return msg

    def step(self, row, add_mc_samples=True):
        return self.msg_from_pred(self.add_xx_to_row_hist_and_make_prediction(self.parse_row_to_xx(row)), add_mc_samples)
