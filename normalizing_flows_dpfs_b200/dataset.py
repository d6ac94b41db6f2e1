"""Disk-tracking dataset loader (reference dataset.py:6-66): an .npz whose `<split>_data` entry is a pickled dict
with start_image, start_state, image, state, q, visible.  Host-side input glue, outside the hot path."""
import os

import numpy as np
import torch
from torch.utils.data import Dataset

_KEYS = ("start_image", "start_state", "image", "state", "q", "visible")


class ToyDiskDataset(Dataset):
    def __init__(self, data_path, filename, datatype="train_data"):
        self.data_path, self.filename = data_path, filename
        split = datatype.split("_")[0]  # train / val / test
        files = sorted(os.path.join(data_path, f) for f in os.listdir(data_path) if f.startswith(filename) and split in f)
        if not files:
            raise FileNotFoundError("no '%s*%s*' file under %s" % (filename, split, data_path))
        blob = dict(np.load(files[0], allow_pickle=True))[datatype].item()  # the reference reads the first file only
        for k in _KEYS:
            setattr(self, k, blob[k])
        self.data_size = len(self.start_image)
        print(self.data_size)

    def __len__(self):
        return self.data_size

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return tuple(getattr(self, k)[idx] for k in _KEYS)
