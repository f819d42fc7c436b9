"""Input pipeline: packed bf16 feature shards, the pinned reader, and the device-side collate tail
against a host restatement of the reference's __getitem__ / collate_fn (vqa_data_memory.py:141-189, 291-396)."""
import numpy as np
import pytest
import torch

from vqacl_b200 import pipeline as P


def _fake_images(n, n_boxes=36, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        w, h = float(torch.randint(300, 640, (1,), generator=g)), float(torch.randint(300, 640, (1,), generator=g))
        feats = torch.relu(torch.randn(n_boxes, 2048, generator=g))
        xy = torch.rand(n_boxes, 2, generator=g) * 0.7
        wh = torch.rand(n_boxes, 2, generator=g) * 0.25 + 0.05
        boxes = torch.cat([xy, xy + wh], dim=1) * torch.tensor([w, h, w, h])
        boxes[0, 2] = w * 1.000001                     # a box touching the border: the reference clamps it to 1.0
        out.append((f"COCO_val2014_{i:012d}", feats.numpy(), boxes.numpy(), w, h))
    return out


def test_pack_read_roundtrip_and_reference_tail(tmp_path):
    imgs = _fake_images(11)
    assert P.pack_features(str(tmp_path), imgs) == 11
    rd = P.PackedFeatureReader(str(tmp_path), pin=False)
    assert len(rd) == 11 and rd.n_boxes == 36 and rd.feat_dim == 2048
    pick = [7, 0, 7, 10, 3]                            # repeated and unordered ids, like a shuffled sampler
    got = rd.gather([imgs[i][0] for i in pick])
    want_f = torch.stack([torch.from_numpy(imgs[i][1]) for i in pick]).bfloat16()          # RNE: the engine's own rounding
    assert got["vis_feats"].dtype == torch.bfloat16 and torch.equal(got["vis_feats"], want_f)
    assert torch.equal(got["boxes_px"], torch.stack([torch.from_numpy(imgs[i][2]) for i in pick]))
    assert torch.equal(got["img_wh"], torch.tensor([[imgs[i][3], imgs[i][4]] for i in pick]))
    # second gather lands in the other staging buffer: the first result is still intact (double buffering)
    first = got["vis_feats"].clone()
    rd.gather([imgs[1][0]] * 5)
    assert torch.equal(got["vis_feats"], first)
    # the host restatement of the tail equals the reference's arithmetic written out literally
    boxes, cate, ques = P.reference_collate_tail(got["boxes_px"], got["img_wh"], [3, 79, 1, 3, 50], [0, 9, 2, 2, 1])
    lit = got["boxes_px"].numpy().copy()
    for k, i in enumerate(pick):
        lit[k][:, (0, 2)] /= imgs[i][3]
        lit[k][:, (1, 3)] /= imgs[i][4]
    lit = np.clip(lit, 0.0, 1.0)
    assert np.array_equal(boxes.numpy(), lit) and boxes.max() <= 1.0
    assert cate.shape == (5, 80) and cate.sum() == 5 and cate[1, 79] == 1 and ques[1, 9] == 1


@pytest.mark.gpu
def test_device_collate_matches_host_tail_and_bf16_features_are_bit_identical(tmp_path):
    import vqacl_b200 as V
    from helpers import O, make_pair
    imgs = _fake_images(8)
    P.pack_features(str(tmp_path), imgs)
    rd = P.PackedFeatureReader(str(tmp_path))
    ids = [imgs[i][0] for i in (5, 1, 2, 7, 0, 3)]
    packed = rd.gather(ids)
    assert packed["vis_feats"].is_pinned()
    b = O.synthetic_batch(6, seed=12, task_id=2)
    cate_ids, ques_ids = b["cate_labels"].argmax(1), b["ques_labels"].argmax(1)
    col = V.DeviceCollator("cuda")
    dev_batch = col(packed, b["input_ids"], b["target_ids"], b["scores"], cate_ids, ques_ids)
    boxes, cate, ques = P.reference_collate_tail(packed["boxes_px"], packed["img_wh"], cate_ids, ques_ids)
    assert torch.equal(dev_batch["boxes"].cpu(), boxes)                       # same IEEE divide + clamp
    assert torch.equal(dev_batch["cate_labels"].cpu(), cate) and torch.equal(dev_batch["ques_labels"].cpu(), ques)
    # the same step from fp32 features through the reference-shaped batch: bit-identical loss, logits and gradients
    _, m = make_pair(layers=2)
    m.train()
    r = m.train_step(dev_batch, 2, 0.5, 0.3)
    loss, logits = r["loss"].detach().clone(), r["logits"].clone()
    r["loss"].backward()
    gWf = m.encoder.visual_embedding.feat_embedding[0].weight.grad.clone()
    V.FusedAdamW(m).zero_grad()
    _, m2 = make_pair(layers=2)
    m2.train()
    ref_batch = dict(b)
    ref_batch["vis_feats"] = torch.stack([torch.from_numpy(imgs[i][1]) for i in (5, 1, 2, 7, 0, 3)])      # fp32, as the reference collates
    ref_batch["boxes"] = boxes
    r2 = m2.train_step(ref_batch, 2, 0.5, 0.3)
    assert torch.equal(r2["loss"].detach(), loss) and torch.equal(r2["logits"], logits)
    r2["loss"].backward()
    g2 = m2.encoder.visual_embedding.feat_embedding[0].weight.grad
    assert torch.allclose(gWf, g2, rtol=1e-4, atol=1e-6)                       # split-K atomics: order-dependent last bits
