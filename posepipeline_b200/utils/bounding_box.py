"""Drop-in for the crop helpers of ``pose_pipeline/utils/bounding_box.py`` (reference :7-53, :101-194): the person crops the
SMPL wrappers consume (``get_person_dataloader``: 224x224 square crops of every present frame) made by the engine's
fixed-point warp kernel instead of one ``cv2.warpAffine`` per frame on the host.

Kept exactly: ``fix_bb_aspect_ratio`` arithmetic, the three corner points and ``cv2.getAffineTransform`` on float32 points
(:45-48; host, a 6x6 solve per frame), the BGR->RGB conversion before the crop (:131), ToTensor + Normalize(ImageNet mean /
std) (:110-116), the skipped absent frames and the returned ``(frame_ids, dataloader, bboxes)`` (:194).  The crop pixels are
bit-identical to ``cv2.warpAffine(INTER_LINEAR)`` (tests/test_gpu_utils.py).
"""
from __future__ import annotations

import os

import cv2
import numpy as np


def fix_bb_aspect_ratio(bbox, dilate=1.2, ratio=1.0):
    """Inflates a bounding box (TLHW) to the desired aspect ratio (width / height); reference :7-29."""
    center = bbox[:2] + bbox[2:] / 2.0
    hw = bbox[2:]
    if hw[0] / hw[1] < ratio:
        hw = np.array([hw[1] * ratio, hw[1]])
    else:
        hw = np.array([hw[0], hw[0] / ratio])
    hw = hw * dilate
    return np.concatenate([center - hw / 2, hw], axis=0)


def crop_transform(bbox, target_size=(288, 384), dilate=1.2):
    """-> (2x3 float64 matrix, corrected bbox): reference :43-48."""
    bbox = fix_bb_aspect_ratio(bbox, ratio=target_size[0] / target_size[1], dilate=dilate)
    src = np.asarray([[bbox[0], bbox[1]], [bbox[0] + bbox[2], bbox[1] + bbox[3]], [bbox[0], bbox[1] + bbox[3]]])
    dst = np.array([[0, 0], [target_size[0], target_size[1]], [0, target_size[1]]])
    return cv2.getAffineTransform(np.float32(src), np.float32(dst)), bbox


def _engine():
    from ..wrappers.mmpose import get_engine
    return get_engine()


def crop_image_bbox(image, bbox, target_size=(288, 384), dilate=1.2, engine=None):
    """reference :32-53, one image: -> (cropped image, corrected bbox)."""
    trans, bbox = crop_transform(bbox, target_size, dilate)
    eng = engine or _engine()
    eng.stage_frames(np.ascontiguousarray(image)[None])
    return eng.warp_affine([0], trans, target_size)[0], bbox


def crop_video_person(video, bboxes, present, crop_size=(224, 224), scale=1.0, engine=None, block=32):
    """All present frames of a video -> (frame_ids, RGB uint8 crops (n,h,w,3), bboxes (n,4)): the loop of reference :123-146
    with blocks of frames decoded by the frame source and cropped on the GPU."""
    from .. import frames as F
    eng = engine or _engine()
    n = len(bboxes)
    frame_ids, crops, out_boxes = [], [], []
    reader = F.BlockReader(video, eng, block, 0, n)
    try:
        for blk in reader:
            # should match the length of identified person tracks
            assert blk.complete and blk.n > 0
            idx = [j for j in range(blk.n) if present[blk.first + j]]
            if idx:
                reader.select(blk)
                tb = [crop_transform(np.asarray(bboxes[blk.first + j]), crop_size, scale) for j in idx]
                # the reference crops the RGB image: same pixels as cropping BGR and swapping channels afterwards
                crops.append(eng.warp_affine(idx, np.stack([t for t, _ in tb]), crop_size, swap_rb=True))
                out_boxes.extend(b for _, b in tb)
                frame_ids.extend(blk.first + j for j in idx)
    finally:
        reader.close()
    crops = np.concatenate(crops) if crops else np.zeros((0, crop_size[1], crop_size[0], 3), np.uint8)
    return frame_ids, crops, np.stack(out_boxes, axis=0) if out_boxes else np.zeros((0, 4))


def get_person_dataloader(key, batch_size=32, num_workers=16, crop_size=224, scale=1.0):
    """reference :101-194: -> (frame_ids, DataLoader of normalised (3,h,w) float tensors, bboxes)."""
    import torch
    from torch.utils.data import DataLoader, Dataset
    from pose_pipeline import PersonBbox, Video

    video, bboxes_dj, present_dj = (Video * PersonBbox & key).fetch1("video", "bbox", "present")
    if type(crop_size) == int or len(crop_size) == 1:
        crop_size = (crop_size, crop_size) if type(crop_size) == int else (crop_size[0], crop_size[0])
    try:
        frame_ids, crops, bboxes = crop_video_person(video, bboxes_dj, present_dj, crop_size, scale)
    finally:
        os.remove(video)
    # transforms.ToTensor() + Normalize(mean, std) (:110-116)
    t = torch.from_numpy(crops).permute(0, 3, 1, 2).to(torch.float32).div(255)
    mean = torch.tensor([0.485, 0.456, 0.406], dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], dtype=torch.float32).view(1, 3, 1, 1)
    frames = list((t - mean) / std)

    class Inference(Dataset):
        def __init__(self, frames, bboxes=None):
            self.frames, self.bboxes, self.scale, self.crop_size = frames, bboxes, scale, crop_size

        def __len__(self):
            return len(self.frames)

        def __getitem__(self, idx):
            return self.frames[idx]

    dataloader = DataLoader(Inference(frames, bboxes), batch_size=batch_size, num_workers=0)
    return frame_ids, dataloader, bboxes
