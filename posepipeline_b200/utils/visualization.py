"""Drop-in for ``video_overlay`` / ``draw_keypoints`` of ``pose_pipeline/utils/visualization.py`` (reference :12-90), the QA
renderer behind ``TopDownPersonVideo.make`` (``pipeline.py:1921-1953``).  The drawing calls are the reference's own (cv2 on
RGB frames, so the output frames are identical); what changes is the plumbing: frames are decoded one block ahead by the
frame source's decode thread and encoded by a writer thread, so decode, drawing and encode overlap instead of running in
one loop.
"""
from __future__ import annotations

import os
import queue
import shutil
import subprocess
import tempfile
import threading

import cv2
import numpy as np


def draw_keypoints(image, keypoints, radius=10, threshold=0.2, color=(255, 255, 255), border_color=(0, 0, 0)):
    """Draw the keypoints on an image (reference :79-90)."""
    image = image.copy()
    keypoints = keypoints.copy()
    keypoints[..., 0] = np.clip(keypoints[..., 0], 0, image.shape[1])
    keypoints[..., 1] = np.clip(keypoints[..., 1], 0, image.shape[0])
    for i in range(keypoints.shape[0]):
        if keypoints[i, -1] > threshold:
            cv2.circle(image, (int(keypoints[i, 0]), int(keypoints[i, 1])), radius, border_color, -1)
            if radius > 2:
                cv2.circle(image, (int(keypoints[i, 0]), int(keypoints[i, 1])), radius - 2, color, -1)
    return image


def video_overlay(video, output_name, callback, downsample=4, codec="MP4V", blur_faces=False, compress=True, bitrate="5M",
                  max_frames=None):
    """Process a video and create overlay image (reference :12-76; same arguments, same frames)."""
    from .. import frames as F
    if blur_faces:
        raise NotImplementedError("blur_faces needs the reference's FaceBlur (facenet wrapper); out of the hot-path scope")
    cap = cv2.VideoCapture(video)
    total_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    h, w = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    fps = cap.get(cv2.CAP_PROP_FPS)
    cap.release()
    output_size = (int(w / downsample), int(h / downsample))
    out = cv2.VideoWriter(output_name, cv2.VideoWriter_fourcc(*codec), fps, output_size)
    if max_frames:
        total_frames = max_frames

    q: "queue.Queue" = queue.Queue(maxsize=64)

    def writer():
        while True:
            f = q.get()
            if f is None:
                return
            out.write(f)
    wt = threading.Thread(target=writer, daemon=True)
    wt.start()
    reader = F.BlockReader(video, None, 16, 0, total_frames)
    try:
        for blk in reader:
            for j in range(blk.n):
                # process image in RGB format
                frame = cv2.cvtColor(blk.frames[j], cv2.COLOR_BGR2RGB)
                out_frame = callback(frame, blk.first + j)
                # move back to BGR format and write to movie
                out_frame = cv2.cvtColor(out_frame, cv2.COLOR_RGB2BGR)
                q.put(cv2.resize(out_frame, output_size))
            if not blk.complete:
                break
    finally:
        reader.close()
        q.put(None)
        wt.join()
        out.release()

    if compress:
        fd, temp = tempfile.mkstemp(suffix=".mp4")
        subprocess.run(["ffmpeg", "-y", "-i", output_name, "-c:v", "libx264", "-b:v", bitrate, temp])
        os.close(fd)
        shutil.move(temp, output_name)
