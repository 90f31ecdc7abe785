"""Dataloader -> device staging: the reference's per-item CPU preprocessing
(dataloader/vidbase.py:52-193) as two CUDA kernels over a whole batch of decoded frame pairs."""
from .pfm import readPFM
from .staging import DeviceStager, ResidentVideo, pair_frame_indices, stage_pairs
from .vidbase import BaseDataset, collate_raw
