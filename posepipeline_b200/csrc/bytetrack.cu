// ByteTrack association on the host: the sequential half of `inference_mot` that the reference runs
// through mmtrack's Python ByteTracker for every frame (reference call site pose_pipeline/wrappers/mmtrack.py:45; tracker
// configuration 3rdparty/mmtracking/mot/bytetrack/bytetrack_yolox_x_crowdhuman_mot17-private-half.py:21-28:
// obj_score_thrs high 0.6 / low 0.1, init_track_thr 0.7, weight_iou_with_det_scores, match_iou_thrs high 0.1 / low 0.5 /
// tentative 0.3, num_frames_retain 30; motion = the DeepSORT constant-velocity Kalman filter on (cx, cy, aspect, h)).
//
// Tens of boxes per frame and a strict frame order: this is microseconds of scalar work, so it stays on the host next to
// the C ABI ("tiny; C++ host Jonker-Volgenant") while the detector batches frames on the GPU.  The assignment
// solves the same extended cost matrix lap.lapjv(extend_cost=True, cost_limit=...) builds, so unless the optimum is tied
// the matches are identical.  IoU arithmetic is float32 like the torch tensors of the reference; the filter is float64.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include "../../include/poseengine.h"
#include "engine_internal.h"

namespace {

struct Track {
  double mean[8];
  double cov[8][8];
  float last_box[5];
  int last_frame = -1;
  int n_boxes = 0;
  bool tentative = false;
};

// ---- Kalman filter (mmtrack/models/motion/kalman_filter.py == deep_sort kalman_filter): x = (cx, cy, a, h, v...)
const double STD_POS = 1.0 / 20, STD_VEL = 1.0 / 160;

void kf_initiate(const double m[4], Track& t) {
  for (int i = 0; i < 4; ++i) { t.mean[i] = m[i]; t.mean[4 + i] = 0; }
  const double h = m[3];
  const double std[8] = {2 * STD_POS * h, 2 * STD_POS * h, 1e-2, 2 * STD_POS * h, 10 * STD_VEL * h, 10 * STD_VEL * h, 1e-5, 10 * STD_VEL * h};
  std::memset(t.cov, 0, sizeof t.cov);
  for (int i = 0; i < 8; ++i) t.cov[i][i] = std[i] * std[i];
}

void kf_predict(Track& t) {
  const double h = t.mean[3];
  const double std[8] = {STD_POS * h, STD_POS * h, 1e-2, STD_POS * h, STD_VEL * h, STD_VEL * h, 1e-5, STD_VEL * h};
  // F = [[I, I], [0, I]]
  for (int i = 0; i < 4; ++i) t.mean[i] += t.mean[4 + i];
  double FP[8][8], P[8][8];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) FP[i][j] = t.cov[i][j] + (i < 4 ? t.cov[i + 4][j] : 0.0);
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) P[i][j] = FP[i][j] + (j < 4 ? FP[i][j + 4] : 0.0);
  for (int i = 0; i < 8; ++i) P[i][i] += std[i] * std[i];
  std::memcpy(t.cov, P, sizeof P);
}

void kf_update(Track& t, const double z[4]) {
  const double h = t.mean[3];
  const double std[4] = {STD_POS * h, STD_POS * h, 1e-1, STD_POS * h};
  double S[4][4], L[4][4] = {{0}};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) S[i][j] = t.cov[i][j] + (i == j ? std[i] * std[i] : 0.0);
  for (int i = 0; i < 4; ++i)           // Cholesky, lower
    for (int j = 0; j <= i; ++j) {
      double s = S[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = (i == j) ? std::sqrt(s) : s / L[j][j];
    }
  // K = P H^T S^-1  (8x4): solve S K^T = (P H^T)^T row by row
  double K[8][4];
  for (int r = 0; r < 8; ++r) {
    double y[4], x[4];
    for (int i = 0; i < 4; ++i) { double s = t.cov[r][i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
    for (int i = 3; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 4; ++k) s -= L[k][i] * x[k]; x[i] = s / L[i][i]; }
    for (int i = 0; i < 4; ++i) K[r][i] = x[i];
  }
  double innov[4];
  for (int i = 0; i < 4; ++i) innov[i] = z[i] - t.mean[i];
  for (int r = 0; r < 8; ++r) { double s = 0; for (int i = 0; i < 4; ++i) s += innov[i] * K[r][i]; t.mean[r] += s; }
  double KS[8][4], P[8][8];
  for (int r = 0; r < 8; ++r)
    for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += K[r][k] * S[k][j]; KS[r][j] = s; }
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) { double s = 0; for (int k = 0; k < 4; ++k) s += KS[r][k] * K[c][k]; P[r][c] = t.cov[r][c] - s; }
  std::memcpy(t.cov, P, sizeof P);
}

void box_to_xyah(const float* b, double out[4]) {       // bbox_xyxy_to_cxcyah in float32, then widened
  const float cx = (b[2] + b[0]) / 2, cy = (b[3] + b[1]) / 2, w = b[2] - b[0], h = b[3] - b[1];
  out[0] = cx; out[1] = cy; out[2] = w / h; out[3] = h;
}

// ---- linear assignment on lap.lapjv's extended matrix: rows n, cols m, unmatched cost `limit`/2 each side
// (shortest augmenting path / Hungarian with potentials, float64).  row_to_col[i] in [0,m) or -1.
void lap_extended(const std::vector<double>& cost, int n, int m, double limit, std::vector<int>& row_to_col, std::vector<int>& col_to_row) {
  row_to_col.assign(n, -1);
  col_to_row.assign(m, -1);
  if (n == 0 || m == 0) return;
  const int N = n + m;
  auto C = [&](int i, int j) -> double {
    if (i < n && j < m) return cost[(size_t)i * m + j];
    if (i >= n && j >= m) return 0.0;
    return limit / 2;
  };
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(N + 1, 0), v(N + 1, 0), minv(N + 1);
  std::vector<int> p(N + 1, 0), way(N + 1, 0);
  std::vector<char> used(N + 1);
  for (int i = 1; i <= N; ++i) {
    p[0] = i;
    int j0 = 0;
    std::fill(minv.begin(), minv.end(), INF);
    std::fill(used.begin(), used.end(), 0);
    do {
      used[j0] = 1;
      const int i0 = p[j0];
      double delta = INF;
      int j1 = 0;
      for (int j = 1; j <= N; ++j) {
        if (used[j]) continue;
        const double cur = C(i0 - 1, j - 1) - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      for (int j = 0; j <= N; ++j) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
    } while (p[j0] != 0);
    do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
  }
  for (int j = 1; j <= m; ++j) {
    const int i = p[j] - 1;
    if (i >= 0 && i < n) { row_to_col[i] = j - 1; col_to_row[j - 1] = i; }
  }
}

}  // namespace

struct pe_bytetrack {
  float high = 0.6f, low = 0.1f, init_thr = 0.7f;
  float iou_high = 0.1f, iou_low = 0.5f, iou_tentative = 0.3f;
  int weight_iou = 1, num_tentatives = 3, retain = 30;
  std::map<int64_t, Track> tracks;    // ordered by id == insertion order (ids only grow), like the reference's dict
  int64_t num_tracks = 0;
};

extern "C" int pe_bytetrack_create(const float* cfg, int32_t n_cfg, pe_bytetrack** out) {
  if (!out || (cfg && n_cfg != 9)) return pe_set_error(PE_ERR_INVALID, "pe_bytetrack_create: cfg must be NULL or 9 floats");
  pe_bytetrack* t = new pe_bytetrack();
  if (cfg) {
    t->high = cfg[0]; t->low = cfg[1]; t->init_thr = cfg[2]; t->iou_high = cfg[3]; t->iou_low = cfg[4]; t->iou_tentative = cfg[5];
    t->weight_iou = cfg[6] != 0; t->num_tentatives = (int)cfg[7]; t->retain = (int)cfg[8];
  }
  *out = t;
  return PE_OK;
}

extern "C" int pe_bytetrack_destroy(pe_bytetrack* t) { delete t; return PE_OK; }

extern "C" int pe_bytetrack_reset(pe_bytetrack* t) {
  if (!t) return pe_set_error(PE_ERR_INVALID, "tracker is NULL");
  t->tracks.clear();
  t->num_tracks = 0;
  return PE_OK;
}

// IoU(track box from the filter state, detection) x optional detection score -> cost = 1 - iou, float32 like the reference
static void assign(pe_bytetrack* t, const std::vector<int64_t>& ids, const std::vector<const float*>& dets, bool weight, float thr,
                   std::vector<int>& row, std::vector<int>& col) {
  const int n = (int)ids.size(), m = (int)dets.size();
  row.assign(n, -1);
  col.assign(m, -1);
  if (n == 0 || m == 0) return;
  std::vector<double> cost((size_t)n * m);
  for (int i = 0; i < n; ++i) {
    const Track& tr = t->tracks[ids[i]];
    // bbox_cxcyah_to_xyxy on a float32 tensor
    const float cx = (float)tr.mean[0], cy = (float)tr.mean[1], a = (float)tr.mean[2], h = (float)tr.mean[3];
    const float w = a * h;
    const float x1 = cx - w / 2.0f, y1 = cy - h / 2.0f, x2 = cx + w / 2.0f, y2 = cy + h / 2.0f;
    const float area1 = (x2 - x1) * (y2 - y1);
    for (int j = 0; j < m; ++j) {
      const float* d = dets[j];
      const float area2 = (d[2] - d[0]) * (d[3] - d[1]);
      const float iw = std::max(std::min(x2, d[2]) - std::max(x1, d[0]), 0.0f);
      const float ih = std::max(std::min(y2, d[3]) - std::max(y1, d[1]), 0.0f);
      const float ov = iw * ih;
      const float uni = std::max(area1 + area2 - ov, 1e-6f);
      float iou = ov / uni;
      if (weight) iou *= d[4];
      cost[(size_t)i * m + j] = (double)(1.0f - iou);
    }
  }
  lap_extended(cost, n, m, (double)(1.0f - thr), row, col);
}

extern "C" int pe_bytetrack_update(pe_bytetrack* t, int32_t frame_id, const float* dets, int32_t n, double* out_rows, int32_t cap,
                                   int32_t* n_out) {
  if (!t || (!dets && n > 0) || n < 0 || !n_out) return pe_set_error(PE_ERR_INVALID, "bad argument to pe_bytetrack_update");
  if (frame_id == 0) { t->tracks.clear(); t->num_tracks = 0; }          // ByteTrack.simple_test: frame 0 resets the tracker
  std::vector<const float*> boxes;
  std::vector<int64_t> ids;
  if (t->tracks.empty() || n == 0) {
    for (int i = 0; i < n; ++i)
      if (dets[5 * i + 4] > t->init_thr) { boxes.push_back(dets + 5 * i); ids.push_back(t->num_tracks++); }
  } else {
    std::vector<const float*> first, second;
    for (int i = 0; i < n; ++i) {
      const float s = dets[5 * i + 4];
      if (s > t->high) first.push_back(dets + 5 * i);
      else if (s > t->low) second.push_back(dets + 5 * i);
    }
    std::vector<int64_t> confirmed, unconfirmed;
    for (auto& kv : t->tracks) (kv.second.tentative ? unconfirmed : confirmed).push_back(kv.first);
    // 1. predict
    for (int64_t id : confirmed) {
      Track& tr = t->tracks[id];
      if (tr.last_frame != frame_id - 1) tr.mean[7] = 0;
      kf_predict(tr);
    }
    // 2. confirmed tracks <-> high-score detections
    std::vector<int> row1, col1;
    assign(t, confirmed, first, t->weight_iou, t->iou_high, row1, col1);
    std::vector<int64_t> first_ids(first.size(), -1);
    for (size_t j = 0; j < first.size(); ++j) if (col1[j] > -1) first_ids[j] = confirmed[col1[j]];
    std::vector<const float*> unmatched;
    std::vector<size_t> unmatched_pos;
    for (size_t j = 0; j < first.size(); ++j) if (col1[j] < 0) { unmatched.push_back(first[j]); unmatched_pos.push_back(j); }
    // 3. tentative tracks <-> still unmatched high-score detections
    std::vector<int> row2, col2;
    assign(t, unconfirmed, unmatched, t->weight_iou, t->iou_tentative, row2, col2);
    std::vector<int64_t> unmatched_ids(unmatched.size(), -1);
    for (size_t j = 0; j < unmatched.size(); ++j) if (col2[j] > -1) unmatched_ids[j] = unconfirmed[col2[j]];
    // 4. confirmed tracks unmatched in step 2 and seen in the previous frame <-> low-score detections (plain IoU)
    std::vector<int64_t> remain;
    for (size_t i = 0; i < confirmed.size(); ++i)
      if (row1[i] == -1 && t->tracks[confirmed[i]].last_frame == frame_id - 1) remain.push_back(confirmed[i]);
    std::vector<int> row3, col3;
    assign(t, remain, second, false, t->iou_low, row3, col3);
    // 5. gather: matched first, unmatched first (tentative matches and new), matched second
    for (size_t j = 0; j < first.size(); ++j) if (col1[j] > -1) { boxes.push_back(first[j]); ids.push_back(first_ids[j]); }
    for (size_t j = 0; j < unmatched.size(); ++j) { boxes.push_back(unmatched[j]); ids.push_back(unmatched_ids[j]); }
    for (size_t j = 0; j < second.size(); ++j) if (col3[j] > -1) { boxes.push_back(second[j]); ids.push_back(remain[col3[j]]); }
    // 6. new ids
    for (auto& id : ids) if (id == -1) id = t->num_tracks++;
  }
  // BaseTracker.update: update or initialise each track, then drop invalid ones
  for (size_t k = 0; k < ids.size(); ++k) {
    double z[4];
    box_to_xyah(boxes[k], z);
    auto it = t->tracks.find(ids[k]);
    if (it != t->tracks.end()) {
      Track& tr = it->second;
      std::memcpy(tr.last_box, boxes[k], sizeof tr.last_box);
      tr.last_frame = frame_id;
      ++tr.n_boxes;
      if (tr.tentative && tr.n_boxes >= t->num_tentatives) tr.tentative = false;
      kf_update(tr, z);
    } else {
      Track tr;
      std::memcpy(tr.last_box, boxes[k], sizeof tr.last_box);
      tr.last_frame = frame_id;
      tr.n_boxes = 1;
      tr.tentative = frame_id != 0;
      kf_initiate(z, tr);
      t->tracks[ids[k]] = tr;
    }
  }
  for (auto it = t->tracks.begin(); it != t->tracks.end();) {
    const Track& tr = it->second;
    const bool lost_too_long = frame_id - tr.last_frame >= t->retain;
    const bool tentative_unmatched = tr.tentative && tr.last_frame != frame_id;
    if (lost_too_long || tentative_unmatched) it = t->tracks.erase(it);
    else ++it;
  }
  // rows [id, x1, y1, x2, y2, score] (float64: the reference concatenates int64 ids with float32 boxes)
  *n_out = (int32_t)ids.size();
  if ((int32_t)ids.size() > cap) return pe_set_error(PE_ERR_INVALID, "pe_bytetrack_update: output capacity too small");
  for (size_t k = 0; k < ids.size(); ++k) {
    double* o = out_rows + 6 * k;
    o[0] = (double)ids[k];
    for (int c = 0; c < 5; ++c) o[1 + c] = (double)boxes[k][c];
  }
  return PE_OK;
}
