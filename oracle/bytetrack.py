"""Oracle: mmtrack ``ByteTracker`` + ``KalmanFilter`` as the reference runs them per frame.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED: mmtrack 0.x is an un-vendored, unpinned dependency
(reference ``requirements.txt:9-12``), reached through ``mmtrack.apis.inference_mot`` at
``pose_pipeline/wrappers/mmtrack.py:45``; it cannot be installed here and the reference ships no golden tracks.  Restated
from the published mmtrack 0.14 sources (``mmtrack/models/trackers/byte_tracker.py``, ``base_tracker.py``,
``mmtrack/models/motion/kalman_filter.py``, ``mmtrack/models/mot/byte_track.py``, ``mmtrack/core/track/transforms.py``)
as configured by ``3rdparty/mmtracking/mot/bytetrack/bytetrack_yolox_x_crowdhuman_mot17-private-half.py:21-28``.

Written independently of the C++ tracker (csrc/bytetrack.cu): numpy float32 boxes / IoUs like the reference's torch
tensors, numpy float64 Kalman filter with scipy's Cholesky like the reference's, and
``scipy.optimize.linear_sum_assignment`` on the extended matrix that ``lap.lapjv(extend_cost=True, cost_limit=...)``
builds (same optimum unless it is tied).
"""
import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

CFG = dict(obj_score_thrs=dict(high=0.6, low=0.1), init_track_thr=0.7, weight_iou_with_det_scores=True,
           match_iou_thrs=dict(high=0.1, low=0.5, tentative=0.3), num_tentatives=3, num_frames_retain=30)


class KalmanFilter:
    """mmtrack/models/motion/kalman_filter.py (the DeepSORT filter): state (cx, cy, a, h, vx, vy, va, vh)."""

    def __init__(self):
        ndim, dt = 4, 1.0
        self._motion_mat = np.eye(2 * ndim, 2 * ndim)
        for i in range(ndim):
            self._motion_mat[i, ndim + i] = dt
        self._update_mat = np.eye(ndim, 2 * ndim)
        self._std_weight_position = 1.0 / 20
        self._std_weight_velocity = 1.0 / 160

    def initiate(self, measurement):
        mean_pos = np.asarray(measurement, np.float64)
        mean = np.r_[mean_pos, np.zeros_like(mean_pos)]
        h = mean_pos[3]
        std = [2 * self._std_weight_position * h, 2 * self._std_weight_position * h, 1e-2, 2 * self._std_weight_position * h,
               10 * self._std_weight_velocity * h, 10 * self._std_weight_velocity * h, 1e-5, 10 * self._std_weight_velocity * h]
        return mean, np.diag(np.square(std))

    def predict(self, mean, covariance):
        std_pos = [self._std_weight_position * mean[3], self._std_weight_position * mean[3], 1e-2, self._std_weight_position * mean[3]]
        std_vel = [self._std_weight_velocity * mean[3], self._std_weight_velocity * mean[3], 1e-5, self._std_weight_velocity * mean[3]]
        motion_cov = np.diag(np.square(np.r_[std_pos, std_vel]))
        mean = np.dot(self._motion_mat, mean)
        covariance = np.linalg.multi_dot((self._motion_mat, covariance, self._motion_mat.T)) + motion_cov
        return mean, covariance

    def project(self, mean, covariance):
        std = [self._std_weight_position * mean[3], self._std_weight_position * mean[3], 1e-1, self._std_weight_position * mean[3]]
        innovation_cov = np.diag(np.square(std))
        mean = np.dot(self._update_mat, mean)
        covariance = np.linalg.multi_dot((self._update_mat, covariance, self._update_mat.T))
        return mean, covariance + innovation_cov

    def update(self, mean, covariance, measurement):
        projected_mean, projected_cov = self.project(mean, covariance)
        chol_factor, lower = scipy.linalg.cho_factor(projected_cov, lower=True, check_finite=False)
        kalman_gain = scipy.linalg.cho_solve((chol_factor, lower), np.dot(covariance, self._update_mat.T).T, check_finite=False).T
        innovation = measurement - projected_mean
        new_mean = mean + np.dot(innovation, kalman_gain.T)
        new_covariance = covariance - np.linalg.multi_dot((kalman_gain, projected_cov, kalman_gain.T))
        return new_mean, new_covariance


def bbox_xyxy_to_cxcyah(b):
    b = np.asarray(b, np.float32)
    cx, cy = (b[2] + b[0]) / np.float32(2), (b[3] + b[1]) / np.float32(2)
    w, h = b[2] - b[0], b[3] - b[1]
    return np.array([cx, cy, w / h, h], np.float32)


def bbox_cxcyah_to_xyxy(b):
    b = np.asarray(b, np.float32)
    cx, cy, ratio, h = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    w = ratio * h
    return np.stack([cx - w / np.float32(2.0), cy - h / np.float32(2.0), cx + w / np.float32(2.0), cy + h / np.float32(2.0)], -1)


def bbox_overlaps(a, b, eps=1e-6):
    """mmdet bbox_overlaps(mode='iou') in float32."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    overlap = wh[..., 0] * wh[..., 1]
    union = np.maximum(area1[:, None] + area2[None] - overlap, np.float32(eps))
    return overlap / union


def lapjv_extended(cost, cost_limit):
    """Row/column assignment of lap.lapjv(cost, extend_cost=True, cost_limit=cost_limit): -1 = unmatched."""
    n, m = cost.shape
    ext = np.full((n + m, n + m), cost_limit / 2.0, np.float64)
    ext[n:, m:] = 0
    ext[:n, :m] = cost
    r, c = linear_sum_assignment(ext)
    row, col = np.full(n, -1, np.int64), np.full(m, -1, np.int64)
    for i, j in zip(r, c):
        if i < n and j < m:
            row[i], col[j] = j, i
    return row, col


class ByteTracker:
    def __init__(self, **cfg):
        c = dict(CFG, **cfg)
        self.obj_score_thrs, self.init_track_thr = c["obj_score_thrs"], c["init_track_thr"]
        self.weight_iou_with_det_scores, self.match_iou_thrs = c["weight_iou_with_det_scores"], c["match_iou_thrs"]
        self.num_tentatives, self.num_frames_retain = c["num_tentatives"], c["num_frames_retain"]
        self.kf = KalmanFilter()
        self.reset()

    def reset(self):
        self.num_tracks, self.tracks = 0, {}

    @property
    def confirmed_ids(self):
        return [i for i, t in self.tracks.items() if not t["tentative"]]

    @property
    def unconfirmed_ids(self):
        return [i for i, t in self.tracks.items() if t["tentative"]]

    def assign_ids(self, ids, det_bboxes, weight_iou_with_det_scores, match_iou_thr):
        track_bboxes = np.zeros((0, 4))
        for i in ids:
            track_bboxes = np.concatenate((track_bboxes, self.tracks[i]["mean"][:4][None]), axis=0)
        track_bboxes = bbox_cxcyah_to_xyxy(track_bboxes.astype(np.float32))
        ious = bbox_overlaps(track_bboxes, det_bboxes[:, :4])
        if weight_iou_with_det_scores:
            ious = ious * det_bboxes[:, 4][None]
        dists = (np.float32(1) - ious).astype(np.float32)
        if dists.size > 0:
            return lapjv_extended(dists.astype(np.float64), float(np.float32(1) - np.float32(match_iou_thr)))
        return np.zeros(len(ids), np.int64) - 1, np.zeros(len(det_bboxes), np.int64) - 1

    def track(self, bboxes, frame_id):
        """bboxes (n,5) float32 [x1,y1,x2,y2,score] -> (bboxes, ids) in the reference's output order."""
        bboxes = np.asarray(bboxes, np.float32).reshape(-1, 5)
        if frame_id == 0:                                   # ByteTrack.simple_test
            self.reset()
        if not self.tracks or len(bboxes) == 0:
            bboxes = bboxes[bboxes[:, -1] > self.init_track_thr]
            ids = np.arange(self.num_tracks, self.num_tracks + len(bboxes))
            self.num_tracks += len(bboxes)
        else:
            ids = np.full(len(bboxes), -1, np.int64)
            first_det_inds = bboxes[:, -1] > self.obj_score_thrs["high"]
            first_det_bboxes, first_det_ids = bboxes[first_det_inds], ids[first_det_inds]
            second_det_inds = (~first_det_inds) & (bboxes[:, -1] > self.obj_score_thrs["low"])
            second_det_bboxes, second_det_ids = bboxes[second_det_inds], ids[second_det_inds]
            for i in self.confirmed_ids:
                t = self.tracks[i]
                if t["frame_ids"][-1] != frame_id - 1:
                    t["mean"][7] = 0
                t["mean"], t["covariance"] = self.kf.predict(t["mean"], t["covariance"])
            confirmed, unconfirmed = self.confirmed_ids, self.unconfirmed_ids
            first_match_track_inds, first_match_det_inds = self.assign_ids(confirmed, first_det_bboxes, self.weight_iou_with_det_scores,
                                                                          self.match_iou_thrs["high"])
            valid = first_match_det_inds > -1
            first_det_ids[valid] = np.asarray(confirmed, np.int64)[first_match_det_inds[valid]]
            first_match_det_bboxes, first_match_det_ids = first_det_bboxes[valid], first_det_ids[valid]
            first_unmatch_det_bboxes, first_unmatch_det_ids = first_det_bboxes[~valid], first_det_ids[~valid]
            _, tentative_match_det_inds = self.assign_ids(unconfirmed, first_unmatch_det_bboxes, self.weight_iou_with_det_scores,
                                                          self.match_iou_thrs["tentative"])
            valid = tentative_match_det_inds > -1
            first_unmatch_det_ids[valid] = np.asarray(unconfirmed, np.int64)[tentative_match_det_inds[valid]]
            first_unmatch_track_ids = [i for k, i in enumerate(confirmed)
                                       if first_match_track_inds[k] == -1 and self.tracks[i]["frame_ids"][-1] == frame_id - 1]
            _, second_match_det_inds = self.assign_ids(first_unmatch_track_ids, second_det_bboxes, False, self.match_iou_thrs["low"])
            valid = second_match_det_inds > -1
            second_det_ids[valid] = np.asarray(first_unmatch_track_ids, np.int64)[second_match_det_inds[valid]]
            valid = second_det_ids > -1
            bboxes = np.concatenate((first_match_det_bboxes, first_unmatch_det_bboxes, second_det_bboxes[valid]), axis=0)
            ids = np.concatenate((first_match_det_ids, first_unmatch_det_ids, second_det_ids[valid]), axis=0)
            new = ids == -1
            ids[new] = np.arange(self.num_tracks, self.num_tracks + new.sum())
            self.num_tracks += int(new.sum())
        # BaseTracker.update
        for i, b in zip(ids, bboxes):
            i = int(i)
            z = bbox_xyxy_to_cxcyah(b)
            if i in self.tracks:
                t = self.tracks[i]
                t["bboxes"].append(b)
                t["frame_ids"].append(frame_id)
                if t["tentative"] and len(t["bboxes"]) >= self.num_tentatives:
                    t["tentative"] = False
                t["mean"], t["covariance"] = self.kf.update(t["mean"], t["covariance"], z)
            else:
                mean, cov = self.kf.initiate(z)
                self.tracks[i] = dict(bboxes=[b], frame_ids=[frame_id], tentative=frame_id != 0, mean=mean, covariance=cov)
        for i in [k for k, v in self.tracks.items()
                  if frame_id - v["frame_ids"][-1] >= self.num_frames_retain or (v["tentative"] and v["frame_ids"][-1] != frame_id)]:
            self.tracks.pop(i)
        return bboxes, ids

    def update(self, frame_id, dets):
        """-> rows [id, x1, y1, x2, y2, score] float64, the layout of result["track_bboxes"][0]."""
        b, i = self.track(dets, frame_id)
        return np.concatenate((np.asarray(i, np.int64)[:, None], b), axis=1) if len(i) else np.zeros((0, 6), np.float64)
