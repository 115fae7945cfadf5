"""Data plumbing names the reference's torchext package exports (torchext/dataset.py:5-66), kept so
`from torchext import *` users find them.  Not on the per-pixel hot path."""
import bisect

import numpy as np
import torch.utils.data


class TestSet(object):
    def __init__(self, name, dset, test_frequency=1):
        self.name, self.dset, self.test_frequency = name, dset, test_frequency


class TestSets(list):
    def append(self, name, dset, test_frequency=1):
        list.append(self, TestSet(name, dset, test_frequency))


class MultiDataset(torch.utils.data.Dataset):
    """Concatenation of datasets addressed by one flat index."""

    def __init__(self, *datasets):
        self.current_epoch = 0
        self.datasets = []
        self.cum_n_samples = [0]
        for d in datasets:
            self.append(d)

    def append(self, dataset):
        self.datasets.append(dataset)
        self.cum_n_samples.append(self.cum_n_samples[-1] + len(dataset))

    def dataset_updated(self):
        self.cum_n_samples = [0]
        for d in self.datasets:
            self.cum_n_samples.append(self.cum_n_samples[-1] + len(d))

    def __len__(self):
        return self.cum_n_samples[-1]

    def __getitem__(self, idx):
        didx = bisect.bisect_right(self.cum_n_samples, idx) - 1
        return self.datasets[didx][idx - self.cum_n_samples[didx]]


class BaseDataset(torch.utils.data.Dataset):
    """Per-sample deterministic RNG: training seeds move with the epoch unless fix_seed_per_epoch."""

    def __init__(self, train=True, fix_seed_per_epoch=False):
        self.current_epoch = 0
        self.train = train
        self.fix_seed_per_epoch = fix_seed_per_epoch

    def get_rng(self, idx):
        if not self.train:
            return np.random.RandomState(idx)
        epoch = 0 if self.fix_seed_per_epoch else self.current_epoch
        return np.random.RandomState((epoch + 1) * len(self) + idx)
