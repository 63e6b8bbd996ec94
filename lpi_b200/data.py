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


# ------------------------------------------------------------------------------------------------------------------------------
# Annotation-schema datasets: the reference's COCO continual split, without its file I/O assumptions.
#
# Reference: retrieval/utils/data.py -- `Coco` (:299-382, train) and `CocoEval` (:185-297, eval) read one JSON list each:
#   train annotation : {"image": <file>, "caption": <str>,   "image_id": <id>, "category": <1..12>}
#   eval  annotation : {"image": <file>, "caption": [<str>], "category": <1..12>}
# A learner task t holds the COCO super-categories TASK_ORDER[t] (12-task setting, data.py:233-247 / :326-340); captions go through
# `pre_caption` (:160-183).  Items: train (image, caption, 0, task); eval (image, index, task) + .text/.text_cat/.image/.txt2img/.img2txt.
# Image decoding is the caller's: `image_loader(name) -> float tensor [3, R, R]` (file loader below, or a synthetic one).
# ------------------------------------------------------------------------------------------------------------------------------
import json
import re
from typing import Callable, Dict, Optional, Sequence, Union

TASK_ORDER = (11, 6, 3, 10, 5, 12, 7, 9, 2, 8, 4, 1)       # task index -> COCO super-category id (data.py:233-247)

_PUNCT = re.compile(r"([,.'!?\"()*#:;~])")
_SPACES = re.compile(r"\s{2,}")


def task_of_category(category: int) -> int:
    """Learner task that owns a super-category; 0 when the category is in no task (the reference's fall-through, data.py:289-292)."""
    return TASK_ORDER.index(category) if category in TASK_ORDER else 0


def pre_caption(caption: str, max_words: int) -> str:
    """data.py:160-183: lower-case, drop , . ' ! ? " ( ) * # : ; ~, '-' and '/' become spaces, '<person>' -> 'person', runs of
    whitespace collapse, at most `max_words` words; an empty result raises ValueError."""
    c = _PUNCT.sub("", caption.lower()).replace("-", " ").replace("/", " ").replace("<person>", "person")
    c = _SPACES.sub(" ", c).rstrip("\n").strip(" ")
    words = c.split(" ")
    if len(words) > max_words:
        c = " ".join(words[:max_words])
    if not len(c):
        raise ValueError("pre_caption yields invalid text")
    return c


def _load_annotations(ann) -> list:
    if isinstance(ann, (str, bytes)):
        with open(ann, "r") as f:
            return json.load(f)
    return list(ann)


def file_image_loader(image_root: str, resolution: int = 224) -> Callable[[str], torch.Tensor]:
    """Deterministic evaluation transform of the reference (Resize(256) -> CenterCrop(224) -> ToTensor -> Normalize, data.py:194-201).
    (The reference's CocoEval DEFAULT transform is the random train augmentation -- quirk C17; loaders there are built with test_trsf.)"""
    import os

    from PIL import Image
    from torchvision import transforms

    tf = transforms.Compose([transforms.Resize(256 * resolution // 224), transforms.CenterCrop(resolution), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    return lambda name: tf(Image.open(os.path.join(image_root, name)).convert("RGB"))


def synthetic_image_loader(res: int = 224, seed: int = 0) -> Callable[[str], torch.Tensor]:
    """Deterministic stand-in for image files: the tensor is a function of the file NAME (there is no dataset on the box)."""
    import zlib

    def load(name: str) -> torch.Tensor:
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
        return torch.randn(3, res, res, generator=g)
    return load


class AnnotationTrainSet(Dataset):
    """`Coco` (data.py:299-382): the annotations whose category belongs to one of `tasks`, in file order (+ an optional replay list)."""

    def __init__(self, ann: Union[str, Sequence[dict]], image_loader: Callable[[str], torch.Tensor], tasks: Sequence[int] = (0,),
                 max_words: int = 30, prompt: str = "", replay_list: Sequence[dict] = ()):
        cats = {TASK_ORDER[t] for t in tasks}
        self.image_loader, self.max_words, self.prompt = image_loader, max_words, prompt
        self.img_ids: Dict = {}
        self.annotation = []
        for a in _load_annotations(ann):
            if a["category"] in cats:
                self.img_ids.setdefault(a["image_id"], len(self.img_ids))
                self.annotation.append(a)
        self.annotation += list(replay_list)

    def __len__(self):
        return len(self.annotation)

    def __getitem__(self, i):
        a = self.annotation[i]
        return self.image_loader(a["image"]), self.prompt + pre_caption(a["caption"], self.max_words), 0, task_of_category(a["category"])


class AnnotationEvalSet(Dataset):
    """`CocoEval` (data.py:185-297): images of the selected tasks with ALL their captions; text ids are assigned in file order."""

    def __init__(self, ann: Union[str, Sequence[dict]], image_loader: Callable[[str], torch.Tensor], tasks: Sequence[int] = (0,),
                 max_words: int = 30):
        cats = {TASK_ORDER[t] for t in tasks}
        self.image_loader = image_loader
        self.ann = [a for a in _load_annotations(ann) if a["category"] in cats]
        self.text, self.text_cat, self.image, self.txt2img, self.img2txt, self.cat = [], [], [], {}, {}, []
        for img_id, a in enumerate(self.ann):
            self.image.append(a["image"])
            self.cat.append(task_of_category(a["category"]))
            self.img2txt[img_id] = []
            for caption in a["caption"]:
                self.img2txt[img_id].append(len(self.text))
                self.txt2img[len(self.text)] = img_id
                self.text.append(pre_caption(caption, max_words))
                self.text_cat.append(self.cat[-1])

    def __len__(self):
        return len(self.ann)

    def __getitem__(self, i):
        return self.image_loader(self.ann[i]["image"]), i, self.cat[i]


def make_synthetic_annotations(n_tasks: int = 5, n_train_per_task: int = 64, n_eval_images_per_task: int = 20, caps_per_image: int = 5,
                               seed: int = 0):
    """(train annotations, eval annotations) in the reference's JSON schema for the first `n_tasks` tasks of TASK_ORDER, interleaved
    across categories the way a real annotation file is (so the task filter has something to do)."""
    g = torch.Generator().manual_seed(6151 * seed + 3)
    train, evals = [], []
    for k in range(n_train_per_task):
        for t in range(n_tasks):
            train.append({"image": f"train/t{t}_{k:05d}.jpg", "caption": _caption(g, t).capitalize() + ".", "image_id": f"coco_{t}_{k}",
                          "category": TASK_ORDER[t]})
    for k in range(n_eval_images_per_task):
        for t in range(n_tasks):
            evals.append({"image": f"val/t{t}_{k:05d}.jpg", "caption": [_caption(g, t).capitalize() + "." for _ in range(caps_per_image)],
                          "category": TASK_ORDER[t]})
    return train, evals


def make_annotation_task_loaders(train_ann, eval_ann, n_tasks: int, batch_size: int = 64, eval_batch_size: int = 128,
                                 image_loader: Optional[Callable[[str], torch.Tensor]] = None):
    """[(train loader of task t, eval loader over tasks 0..t)] from annotation lists / files -- what `SPrompts.incremental_train` consumes,
    built the way the reference's `incremental_train` builds them per task (sprompt.py:154-172)."""
    image_loader = image_loader or synthetic_image_loader()
    return [(DataLoader(AnnotationTrainSet(train_ann, image_loader, tasks=[t]), batch_size=batch_size, shuffle=False),
             DataLoader(AnnotationEvalSet(eval_ann, image_loader, tasks=list(range(t + 1))), batch_size=eval_batch_size, shuffle=False))
            for t in range(n_tasks)]
