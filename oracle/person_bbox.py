"""Oracle: ``PersonBbox.make`` (reference ``pose_pipeline/pipeline.py:656-687``), restated with pandas.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PINNED: tests/golden/person_bbox_*.npz are produced
by executing the reference's own ``make`` body (tests/golden/make_golden.py) and this restatement
must reproduce them bit for bit.

pandas >= 2.1 removed ``fillna(method=...)`` (``pipeline.py:680-681`` crashes on this container's
pandas 3); ``bfill(limit=2)`` / ``ffill(limit=2)`` are the same operation.
"""
import numpy as np
import pandas as pd


def person_bbox(tracks, keep_tracks):
    """tracks: list (frames) of list of dicts {track_id, tlhw,...}; -> (bbox (N,4) float64 w/ NaN rows, present (N,) bool)."""
    def process_timestamp(track_timestep):                       # :662-667
        valid = [t for t in track_timestep if t["track_id"] in keep_tracks]
        if len(valid) == 1:
            return {"present": True, "bbox": valid[0]["tlhw"]}
        return {"present": False, "bbox": [0.0, 0.0, 0.0, 0.0]}

    LD = [process_timestamp(t) for t in tracks]                  # :669-671
    present = np.array([d["present"] for d in LD])               # :674
    bbox = np.array([d["bbox"] for d in LD])                     # :675
    df = pd.DataFrame(bbox)                                      # :678
    df.iloc[~present] = np.nan                                   # :679
    df = df.bfill(axis=0, limit=2)                               # :680
    df = df.ffill(axis=0, limit=2)                               # :681
    return df.values, ~df.isna().any(axis=1).values              # :684-685
