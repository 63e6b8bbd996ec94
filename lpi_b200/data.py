"""In-memory synthetic continual-retrieval tasks in the reference's item formats (the COCO file datasets of
retrieval/utils/data.py are outside the hot path; there is no dataset on the box).

Train items: (image [3,224,224], caption: str, 0, task)            -- utils/data.py `Coco.__getitem__`, consumed at sprompt.py:300
Eval items : (image, index, task); the dataset exposes .text, .text_cat, .image, .txt2img, .img2txt
                                                                   -- `CocoEval`, consumed at sprompt.py:441-451, 456, 546
Images are a task-specific mean image plus noise, captions are drawn from task-specific word subsets, so the K-Means task
keys separate the tasks."""
from __future__ import annotations

from typing import List, Tuple

import torch
from torch.utils.data import DataLoader, Dataset

from .synthetic import _WORDS


def _task_mean(task: int, res: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(50021 + task)
    return torch.randn(3, res // 16, res // 16, generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2) * 0.8


def _caption(g: torch.Generator, task: int, n_tasks_words: int = 12) -> str:
    per = len(_WORDS) // n_tasks_words
    sub = _WORDS[(task % n_tasks_words) * per:(task % n_tasks_words + 1) * per] + _WORDS[:6]
    n = int(torch.randint(6, 13, (1,), generator=g))
    return " ".join(sub[int(i)] for i in torch.randint(0, len(sub), (n,), generator=g))


class SyntheticTrainSet(Dataset):
    def __init__(self, task: int, n_pairs: int = 256, res: int = 224, seed: int = 0):
        g = torch.Generator().manual_seed(7001 * seed + 31 * task + 1)
        self.task = task
        self.images = _task_mean(task, res)[None] + torch.randn(n_pairs, 3, res, res, generator=g)
        self.captions = [_caption(g, task) for _ in range(n_pairs)]

    def __len__(self):
        return len(self.captions)

    def __getitem__(self, i):
        return self.images[i], self.captions[i], 0, self.task


class SyntheticEvalSet(Dataset):
    def __init__(self, tasks: List[int], n_images: int = 100, caps_per_image: int = 5, res: int = 224, seed: int = 0):
        self.image, self.text, self.text_cat, self.img2txt, self.txt2img, self.cat = [], [], [], {}, {}, []
        for task in tasks:
            g = torch.Generator().manual_seed(9001 * seed + 37 * task + 5)
            imgs = _task_mean(task, res)[None] + torch.randn(n_images, 3, res, res, generator=g)
            for k in range(n_images):
                idx = len(self.image)
                self.image.append(imgs[k])
                self.cat.append(task)
                self.img2txt[idx] = []
                for _ in range(caps_per_image):
                    self.img2txt[idx].append(len(self.text))
                    self.txt2img[len(self.text)] = idx
                    self.text.append(_caption(g, task))
                    self.text_cat.append(task)

    def __len__(self):
        return len(self.image)

    def __getitem__(self, i):
        return self.image[i], i, self.cat[i]


def make_task_loaders(n_tasks: int = 5, n_train: int = 256, n_eval_images: int = 100, caps_per_image: int = 5, batch_size: int = 64,
                      eval_batch_size: int = 128, res: int = 224, seed: int = 0) -> List[Tuple[DataLoader, DataLoader]]:
    """[(train_loader of task i, test_loader over tasks 0..i)] -- the loaders `SPrompts.incremental_train` consumes."""
    out = []
    for t in range(n_tasks):
        tr = DataLoader(SyntheticTrainSet(t, n_train, res, seed), batch_size=batch_size, shuffle=False)
        te = DataLoader(SyntheticEvalSet(list(range(t + 1)), n_eval_images, caps_per_image, res, seed), batch_size=eval_batch_size, shuffle=False)
        out.append((tr, te))
    return out
