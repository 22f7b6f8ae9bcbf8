"""In-memory stand-in for the slice of the reference's DataJoint schema the hot path touches: enough of
``Table & key``, ``fetch1``, ``insert1`` and ``Video.get_robust_reader`` for the wrappers / make() bodies to run without MySQL."""
import os
import shutil
import sys
import tempfile
import types


class _Query:
    def __init__(self, table, rows):
        self.table, self.rows = table, rows

    def __and__(self, key):
        if isinstance(key, dict):
            pk = [k for k in key if any(k in r for r in self.rows)]
            rows = [r for r in self.rows if all(r.get(k) == key[k] for k in pk if k in r)]
        else:
            rows = self.rows
        return _Query(self.table, rows)

    def fetch1(self, *attrs):
        assert len(self.rows) == 1, f"fetch1 on {len(self.rows)} rows of {self.table.__name__}"
        vals = tuple(self.rows[0][a] for a in attrs)
        return vals[0] if len(vals) == 1 else vals

    def __len__(self):
        return len(self.rows)


class _TableMeta(type):
    def __and__(cls, key):
        return _Query(cls, cls.rows) & key

    def __len__(cls):
        return len(cls.rows)


class Table(metaclass=_TableMeta):
    rows = []

    def insert1(self, row, **kw):
        type(self).rows.append(dict(row))

    @classmethod
    def clear(cls):
        cls.rows = []


def make_fake_pose_pipeline(model_data_dir=""):
    """Builds and registers fake ``pose_pipeline`` / ``pose_pipeline.pipeline`` modules; returns the namespace."""
    names = ["Video", "VideoInfo", "TrackingBbox", "PersonBboxValid", "PersonBbox", "TopDownPerson", "LiftingPerson"]
    ns = {n: type(n, (Table,), {"rows": []}) for n in names}

    def get_robust_reader(key, return_cap=True):
        video = (ns["Video"] & key).fetch1("video")
        fd, outfile = tempfile.mkstemp(suffix=".mp4")
        os.close(fd)
        shutil.copy(video, outfile)           # the reference moves DataJoint's fetched copy (pipeline.py:53-56)
        assert not return_cap
        return outfile

    ns["Video"].get_robust_reader = staticmethod(get_robust_reader)
    pkg = types.ModuleType("pose_pipeline")
    pipe = types.ModuleType("pose_pipeline.pipeline")
    wr = types.ModuleType("pose_pipeline.wrappers")
    for k, v in ns.items():
        setattr(pkg, k, v)
        setattr(pipe, k, v)
    pkg.MODEL_DATA_DIR = model_data_dir
    pkg.pipeline, pkg.wrappers = pipe, wr
    sys.modules["pose_pipeline"] = pkg
    sys.modules["pose_pipeline.pipeline"] = pipe
    sys.modules["pose_pipeline.wrappers"] = wr
    return ns


def write_video(path, frames, fps=30):
    import cv2
    h, w = frames[0].shape[:2]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    assert vw.isOpened()
    for f in frames:
        vw.write(f)
    vw.release()


def read_video(path):
    import cv2
    cap = cv2.VideoCapture(path)
    out = []
    while True:
        ret, f = cap.read()
        if not ret:
            break
        out.append(f)
    cap.release()
    return out
