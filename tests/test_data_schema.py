"""SURVEY.md section 8(f) f4 on CPU: the annotation-schema datasets (lpi_b200/data.py) against the reference's own `Coco` / `CocoEval`
(retrieval/utils/data.py:185-382) on the same JSON files, plus `pre_caption` and the 12-task category order."""
import json
import os
import sys

import pytest
import torch

from lpi_b200 import data as D
from oracle import reference_loader as RL

needs_ref = pytest.mark.skipif(not RL.reference_available(), reason="reference tree not mounted")


def test_task_order_and_pre_caption_known_answers():
    assert D.TASK_ORDER == (11, 6, 3, 10, 5, 12, 7, 9, 2, 8, 4, 1) and D.task_of_category(11) == 0 and D.task_of_category(1) == 11
    assert D.task_of_category(99) == 0                                     # the reference's fall-through
    assert D.pre_caption("A man, riding a wave-board/skate;  on   the <person>'s LAWN!\n", 30) == "a man riding a wave board skate on the persons lawn"
    assert D.pre_caption("one two three four", 3) == "one two three"
    with pytest.raises(ValueError):
        D.pre_caption(" .,! ", 30)


def test_synthetic_annotations_flow_through_the_loaders():
    train, evals = D.make_synthetic_annotations(3, 8, 4, 2)
    assert set(train[0]) == {"image", "caption", "image_id", "category"} and set(evals[0]) == {"image", "caption", "category"}
    loaders = D.make_annotation_task_loaders(train, evals, 3, batch_size=4, eval_batch_size=4, image_loader=D.synthetic_image_loader(32))
    assert len(loaders) == 3
    for t, (tr, te) in enumerate(loaders):
        imgs, caps, zero, task = next(iter(tr))
        assert imgs.shape == (4, 3, 32, 32) and len(caps) == 4 and int(zero[0]) == 0 and set(task.tolist()) == {t}
        ds = te.dataset
        assert len(ds) == 4 * (t + 1) and len(ds.text) == 2 * len(ds) and set(ds.text_cat) == set(range(t + 1))
        assert all(ds.txt2img[c] == i for i, cs in ds.img2txt.items() for c in cs)
        img, idx, cat = ds[len(ds) - 1]
        assert idx == len(ds) - 1 and cat == ds.cat[idx]
    a = D.synthetic_image_loader(32)("x.jpg")
    assert torch.equal(a, D.synthetic_image_loader(32)("x.jpg")) and not torch.equal(a, D.synthetic_image_loader(32)("y.jpg"))


@needs_ref
def test_annotation_sets_match_reference_coco_classes(tmp_path):
    RL.load_reference()                                                    # puts the reference's retrieval/ on sys.path (+ stubs)
    import utils.data as ref_data

    train, evals = D.make_synthetic_annotations(4, 6, 3, 3, seed=1)
    train.append({"image": "train/odd.jpg", "caption": "Some-thing (odd): <person> #1 / two", "image_id": "coco_0_0", "category": 6})
    tf, ef = str(tmp_path / "train.json"), str(tmp_path / "val.json")
    json.dump(train, open(tf, "w")); json.dump(evals, open(ef, "w"))
    for tasks in ([0], [1], [0, 1, 2], [3, 1]):
        ref = ref_data.CocoEval(image_root=str(tmp_path), ann_file=ef, tasks=tasks)
        mine = D.AnnotationEvalSet(ef, D.synthetic_image_loader(32), tasks=tasks)
        assert len(mine) == len(ref) and mine.text == ref.text and mine.text_cat == ref.text_cat and mine.image == ref.image
        assert mine.txt2img == ref.txt2img and mine.img2txt == ref.img2txt
        ref_tr = ref_data.Coco(image_root=str(tmp_path), ann_file=tf, tasks=tasks, prompt="")
        mine_tr = D.AnnotationTrainSet(tf, D.synthetic_image_loader(32), tasks=tasks)
        assert mine_tr.annotation == ref_tr.annotation and mine_tr.img_ids == ref_tr.img_ids and len(mine_tr) == len(ref_tr)
        for i in range(len(mine_tr)):                                      # caption + task of every item (the image needs files: not compared)
            a = ref_tr.annotation[i]
            want_task = next((z for z in range(len(ref_tr.tasks)) if a["category"] in ref_tr.tasks[z]), 0)
            _, cap, zero, task = mine_tr[i]
            assert cap == ref_data.pre_caption(a["caption"], 30) and zero == 0 and task == want_task
    assert [t[0] for t in ref.tasks] == list(D.TASK_ORDER)


@needs_ref
def test_file_image_loader_matches_reference_eval_transform(tmp_path):
    RL.load_reference()
    import numpy as np
    import utils.data as ref_data
    from PIL import Image
    from torchvision import transforms

    rng = np.random.default_rng(0)
    Image.fromarray(rng.integers(0, 255, (300, 400, 3), dtype=np.uint8)).save(tmp_path / "a.png")
    evals = [{"image": "a.png", "caption": ["a b c"], "category": 11}]
    ef = str(tmp_path / "val.json")
    json.dump(evals, open(ef, "w"))
    tf = transforms.Compose([*ref_data.CocoEval.test_trsf, *ref_data.CocoEval.common_trsf])
    ref = ref_data.CocoEval(transform=tf, image_root=str(tmp_path), ann_file=ef, tasks=[0])
    mine = D.AnnotationEvalSet(ef, D.file_image_loader(str(tmp_path)), tasks=[0])
    (ri, rx, rc), (mi, mx, mc) = ref[0], mine[0]
    assert torch.equal(ri, mi) and rx == mx and rc == mc
