"""The per-stream estimation loop behind the reference's ``Estimator`` API (``estimate/estimator.py:17-218``).

Constructor arguments, methods, properties, thread / queue behaviour and error behaviour follow the reference so
the socket listeners, UDP publishers and CSV recorders around it keep working unchanged.  The three per-frame
calls (``parse_row_to_xx`` -> ``add_xx_to_row_hist_and_make_prediction`` -> ``msg_from_pred``, estimator.py:174-176)
each run their CUDA stage when called on their own; ``processing_loop`` uses the fused single-launch-sequence
path of ``BatchedEstimator`` with one stream (one H2D copy of the row, one D2H copy of the message).
"""
import logging
import queue
import threading
from datetime import datetime

import numpy as np
import torch

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.data_types.bone_map import BoneMap
from arm_pose_estimation_b200.estimate import estimate_joints
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator
from arm_pose_estimation_b200.utility import data_stats
from arm_pose_estimation_b200.utility.names import NNS_INPUTS, NNS_TARGETS


class Estimator:
    # set by the NN subclasses: which parse_row_to_xx the feature kernel follows, and the wire layout of a row
    _kind = None
    _layout = None

    def __init__(self,
                 x_inputs: NNS_INPUTS,
                 y_targets: NNS_TARGETS,
                 normalize: bool = True,
                 smooth: int = 1,
                 seq_len: int = 1,
                 add_mc_samples: bool = True,
                 bonemap: BoneMap = None,
                 tag: str = "Estimator"):
        self.__tag = tag
        self._active = False
        self._y_targets = y_targets
        self._x_inputs = x_inputs
        self._normalize = normalize
        if normalize:
            stats = data_stats.get_norm_stats(x_inputs=self._x_inputs, y_targets=self._y_targets)
            self._xx_m, self._xx_s = stats["xx_m"], stats["xx_s"]
            self._yy_m, self._yy_s = stats["yy_m"], stats["yy_s"]
        self._smooth = max(1, smooth)
        self._smooth_hist = []
        self._last_msg = None
        self._add_mc_samples = add_mc_samples
        self._row_hist = []
        self._sequence_len = max(1, seq_len)
        if bonemap is None:
            self._larm_vec = np.array([-BoneMap.DEFAULT_LARM_LEN, 0, 0])
            self._uarm_vec = np.array([-BoneMap.DEFAULT_UARM_LEN, 0, 0])
            self._uarm_orig = BoneMap.DEFAULT_UARM_ORIG_RH
        else:
            self._larm_vec = np.array([-bonemap.left_lower_arm_length, 0, 0])
            self._uarm_vec = np.array([-bonemap.left_upper_arm_length, 0, 0])
            self._uarm_orig = bonemap.left_upper_arm_origin_rh
        self._body_measurements = np.r_[self._larm_vec, self._uarm_vec, self._uarm_orig][np.newaxis, :]
        self._device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self._bonemap = bonemap
        self._fused = None            # BatchedEstimator with one stream, built on first use

    def set_norm_stats(self, stats: dict):
        """overwrites the default norm stats loaded during the initialization"""
        self._xx_m, self._xx_s = stats["xx_m"], stats["xx_s"]
        self._yy_m, self._yy_s = stats["yy_m"], stats["yy_s"]
        self._fused = None
        logging.info("Replaced norm stats xx m+/-s and yy m+/-s")

    def get_last_msg(self):
        return self._last_msg

    def is_active(self):
        return self._active

    def terminate(self):
        self._active = False

    def reset(self):
        self._active = False
        self._row_hist = []
        self._smooth_hist = []
        if self._fused is not None:
            self._fused.reset()

    # ---- the three per-frame calls, each on its own (estimator.py:93-137) -------------------------------
    def add_xx_to_row_hist_and_make_prediction(self, xx) -> np.array:
        self._row_hist.append(xx)
        while len(self._row_hist) < self._sequence_len:      # first frame: repeat the row (estimator.py:96-97)
            self._row_hist.append(xx)
        while len(self._row_hist) > self._sequence_len:
            del self._row_hist[0]
        xx_hist = np.vstack(self._row_hist)
        if self._normalize:
            xx_hist = (xx_hist - self._xx_m) / self._xx_s
        pred = self.make_prediction_from_row_hist(xx_hist)
        if self._normalize:
            pred = pred * self._yy_s + self._yy_m
        if self._smooth > 1:
            self._smooth_hist.append(pred)
            while len(self._smooth_hist) < self._smooth:
                self._smooth_hist.append(pred)
            while len(self._smooth_hist) > self._smooth:
                del self._smooth_hist[0]
            pred = np.vstack(self._smooth_hist)
        return pred

    def msg_from_pred(self, pred: np.array, add_mc_samples: bool) -> np.array:
        est, msg, _ = estimate_joints.fk_rows(pred, self._body_measurements, self._y_targets, want_msg=True)
        self._last_msg = msg.copy()
        if add_mc_samples:
            msg = list(msg)
            if est.shape[0] > 1:                             # message tail: hand xyz, elbow xyz of every row
                msg += est[:, :6].ravel().tolist()
        return msg

    # ---- fused path: the same three stages as one launch sequence ---------------------------------------
    def _fused_estimator(self):
        if self._fused is None:
            self._fused = self._make_fused()
        return self._fused

    def _make_fused(self):
        raise UserWarning("this estimator has no fused CUDA path")

    def estimate_row(self, row, add_mc_samples=None):
        """One frame of the loop body (estimator.py:174-176) through the fused path; returns the message."""
        add = self._add_mc_samples if add_mc_samples is None else add_mc_samples
        fe = self._fused_estimator()
        out = fe.step_graph(np.asarray(row, dtype=np.float32).reshape(1, 1, -1))     # one CUDA-graph launch per frame
        if int(out.status[0, 0]) != 0:
            raise np.linalg.LinAlgError("degenerate 6D rotation (zero or collinear columns)")
        msg = out.msg[0, 0].astype(np.float64)
        self._last_msg = msg.copy()
        self._last_std = out.std[0, 0].astype(np.float64)
        if add:
            msg = list(msg)
            if fe.S > 1:
                msg += out.samples[0, 0].astype(np.float64).ravel().tolist()
        return msg

    def get_last_std(self):
        """Population std of the per-row hand / elbow positions of the last fused estimate (new)."""
        return getattr(self, "_last_std", None)

    def process_in_thread(self, sensor_q: queue):
        msg_q = queue.Queue()
        t = threading.Thread(target=self.processing_loop, args=(sensor_q, msg_q))
        t.start()
        return msg_q

    def processing_loop(self, sensor_q: queue, msg_q: queue):
        logging.info(f"[{self.__tag}] wearable streaming loop")
        start = datetime.now()
        dat = 0
        self.reset()
        self._active = True
        while self._active:
            try:
                row = sensor_q.get(timeout=2)
                while sensor_q.qsize() > 5:                  # freshness: shed the backlog (estimator.py:160-161)
                    row = sensor_q.get(timeout=2)
            except queue.Empty:
                logging.info(f"[{self.__tag}] no data")
                continue
            now = datetime.now()
            if (now - start).seconds >= 5:
                start = now
                logging.info(f"[{self.__tag}] {dat / 5} Hz")
                dat = 0
            msg = self.estimate_row(row, self._add_mc_samples)
            msg_q.put(msg)
            dat += 1

    def make_prediction_from_row_hist(self, xx_hist: np.array) -> np.array:
        raise NotImplementedError

    def parse_row_to_xx(self, row) -> np.array:
        raise NotImplementedError

    @property
    def sequence_len(self):
        return self._sequence_len

    @property
    def body_measurements(self):
        return self._body_measurements

    @property
    def uarm_orig(self):
        return self._uarm_orig

    @property
    def uarm_vec(self):
        return self._uarm_vec

    @property
    def larm_vec(self):
        return self._larm_vec

    @property
    def device(self):
        return self._device

    @property
    def x_inputs(self):
        return self._x_inputs

    @property
    def y_targets(self):
        return self._y_targets


class _NNEstimator(Estimator):
    """What the three NN estimators share: model loading, the feature kernel per row, the MC forward pass."""
    _xx_dtype = np.float32

    def _init_nn(self, model_hash, smooth, add_mc_samples, monte_carlo_samples, bonemap, tag, philox_seed):
        from arm_pose_estimation_b200.estimate import nn_models
        self._mc_samples = monte_carlo_samples
        self._nn_model, params = nn_models.load_deployed_model_from_hash(hash_str=model_hash)
        self._params = params
        self._philox_seed = self._nn_model.philox_seed if philox_seed is None else int(philox_seed)
        self._nn_model.philox_seed = self._philox_seed
        Estimator.__init__(
            self,
            x_inputs=NNS_INPUTS[params["x_inputs_n"]],
            y_targets=NNS_TARGETS[params["y_targets_n"]],
            smooth=smooth,
            normalize=params["normalize"],
            seq_len=params["sequence_len"],
            add_mc_samples=add_mc_samples,
            tag=tag,
            bonemap=bonemap,
        )

    def parse_row_to_xx(self, row):
        """Calibrated feature row of one raw IMU row (the estimator's ``parse_row_to_xx`` of the reference),
        computed by the stage-1 kernel with ``normalize=0``."""
        if not torch.cuda.is_available():
            raise RuntimeError("arm_pose_estimation_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        ncols = 28 if self._layout == N.LAYOUT_WATCH_ONLY else 55
        r = np.asarray(row, dtype=np.float32).ravel()
        if r.size < ncols:
            raise UserWarning(f"a raw row of this estimator has {ncols} floats, got {r.size}")
        raw = torch.from_numpy(np.ascontiguousarray(r[:ncols])).cuda()
        I = len(self._x_inputs.value)
        out = torch.empty(I, dtype=torch.float32, device="cuda")
        N.check(N.load().ape_features(N.ptr(raw), self._layout, self._kind, None, None, 0, N.ptr(out), 1, 1, 0, None, 1,
                                      N.current_stream_ptr()), "ape_features")
        return out.cpu().numpy().astype(self._xx_dtype)

    def make_prediction_from_row_hist(self, xx_hist, masks=None):
        xx = torch.tensor(np.asarray(xx_hist)[None, :, :], dtype=torch.float32)
        t_preds = self._nn_model.monte_carlo_predictions(x=xx, n_samples=self._mc_samples, masks=masks)
        return t_preds.numpy()[:, -1, :]                     # only the last step of the sequence (watch_only.py:97)

    def _make_fused(self, mask_mode=N.MASK_PHILOX):
        stats = dict(xx_m=self._xx_m, xx_s=self._xx_s, yy_m=self._yy_m, yy_s=self._yy_s) if self._normalize else None
        return BatchedEstimator(
            kind=self._kind, layout=self._layout, state=self._nn_model.state_dict(), seq_len=self._sequence_len,
            y_targets=self._y_targets, stats=stats, n_streams=1, mc_samples=self._mc_samples, smooth=self._smooth,
            dropout=self._nn_model.dropout, bonemap=self._bonemap, frames_per_call=1, emit_samples=True,
            normalize=self._normalize, mask_mode=mask_mode, philox_seed=self._philox_seed)
