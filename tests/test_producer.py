"""The PyTorch input stage (rcvpose_b200/producer.py, BASELINE configs[1]): the trunk of the reference's DenseFCNResNet152 up to
conv7, with the reference's parameter names, against the reference model's own outputs (tests/golden/make_golden_producer.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rcvpose_b200 import producer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


@pytest.fixture(scope="module")
def seeded():
    from make_golden_producer_helpers import seeded_trunk
    return seeded_trunk()


def test_trunk_matches_the_reference_model_outputs(seeded):
    """Same state_dict, same input: seg / radial maps of the reference's network (golden, computed by the reference model itself)
    against RadiusTrunk + conv8 in PyTorch fp32: 1e-4 relative (the same convolutions; only the kernel selection may differ)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "producer_golden.npz"))
    net, x = seeded
    assert abs(float(x.double().sum()) - float(g["x_checksum"])) < 1e-9
    keys = "\n".join("%s %s" % (k, tuple(v.shape)) for k, v in net.state_dict().items())
    assert keys == str(g["keys"]), "parameter names / shapes differ from the reference's DenseFCNResNet152: checkpoints would not load"
    with torch.no_grad():
        seg, rad = net.forward_reference(x.clone())
    for got, want in ((seg.numpy(), g["seg"]), (rad.numpy(), g["radial"])):
        assert got.shape == want.shape == (2, 1, 64, 96)
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    w, b = net.head()
    with torch.no_grad():
        up = net(x.clone())
    assert up.shape == (2, 32, 64, 96) and float(up.min()) >= 0.0            # conv7 + BN + ReLU: what the head kernel consumes
    manual = torch.einsum("nk,bkhw->bnhw", w, up) + b[None, :, None, None]
    assert torch.allclose(manual[:, :1], seg, atol=1e-6) and torch.allclose(manual[:, 1:], rad, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="the reference tree exists in the build container only")
def test_trunk_is_state_dict_compatible_with_the_reference(seeded):
    sys.path.insert(0, "/root/reference")
    from models.fcnresnet import DenseFCNResNet152
    net, x = seeded
    ref = DenseFCNResNet152(3, 2).eval()
    ref.load_state_dict(net.state_dict(), strict=True)          # a reference checkpoint loads the other way round just the same
    net2 = producer.RadiusTrunk().eval()
    net2.load_state_dict(ref.state_dict(), strict=True)
    with torch.no_grad():
        s_ref, r_ref = ref(x.clone())
        s, r = net2.forward_reference(x.clone())
    assert torch.equal(s, s_ref) and torch.equal(r, r_ref)


def test_normalise_rgb_is_the_reference_preprocessing():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(5, 7, 3)).astype(np.uint8)
    t = producer.normalise_rgb(img)
    want = np.array(img, dtype=np.float64)          # AccumulatorSpace.py:143-147
    want /= 255.
    want -= np.array([0.485, 0.456, 0.406])
    want /= np.array([0.229, 0.224, 0.225])
    assert t.shape == (1, 3, 5, 7) and t.dtype == torch.float32
    assert np.array_equal(t[0].numpy(), want.transpose(2, 0, 1).astype(np.float32))
    assert producer.normalise_rgb(np.stack([img, img])).shape == (2, 3, 5, 7)


def test_tail_fold_is_conv7_bn_relu():
    """RadiusTrunk.tail(): conv7's bias and its eval-mode BatchNorm folded into scale / shift (what rcv_conv7_head applies after
    the implicit GEMM) reproduce the module (models/fcnresnet.py:114-116)."""
    torch.manual_seed(2)
    t = producer.RadiusTrunk().eval()
    bn = t.conv7[1]
    with torch.no_grad():
        bn.running_mean.normal_(0.0, 0.2); bn.running_var.uniform_(0.3, 2.0); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0.0, 0.2)
        x = torch.randn((2, 64, 9, 11))
        want = t.conv7(x)
        w7, sc, sh, w8, b8 = t.tail()
        got = torch.relu(F.conv2d(x, w7, padding=1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
        assert w8.shape == (2, 32) and b8.shape == (2,)
        assert torch.allclose(t(torch.zeros((1, 3, 32, 32))), t.conv7(t.forward_up1(torch.zeros((1, 3, 32, 32)))))


_pending = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def _pending_gpu(fn):
    for m in _pending:
        fn = m(fn)
    return fn


@_pending_gpu
def test_producer_stage_feeds_the_fused_head_vote():
    """BASELINE configs[1] plumbing on the GPU: three bf16 trunks -> conv7 activations -> rcv_head_vote_frames, bit-identical to
    rcv_head_1x1 + rcv_vote_frames on the same activations; the bf16 activations agree with the fp32 trunk to bf16 accuracy."""
    from rcvpose_b200 import api, synth
    ctx = api.VoteContext(0, max_items=16, max_points_total=1 << 22, max_grid=400)
    torch.manual_seed(5)
    trunks = [producer.RadiusTrunk() for _ in range(3)]
    stage = producer.ProducerStage(trunks, ctx)
    frames = [synth.config3_frame(f) for f in (41, 42)]
    rgb = torch.rand((2, 3, 480, 640))
    depth = torch.from_numpy(np.stack([f["depth"] for f in frames]).view(np.int16)).cuda()
    K = torch.from_numpy(frames[0]["K"]).cuda()
    mr = torch.full((3,), 1.0e9, dtype=torch.float64, device="cuda")
    up = stage.activations(rgb)
    assert up.shape == (2, 3, 32, 480, 640) and up.dtype == torch.bfloat16 and up.is_contiguous() and bool(torch.isfinite(up.float()).all())
    fused = ctx.head_vote_frames(up, stage.weight, stage.bias, depth, K, max_radii=mr, mask_flags=api.RCV_MASK_MAX_RADIUS | api.RCV_MASK_RADIUS_POSITIVE,
                                 want_radius=True)
    maps = torch.stack([ctx.head_1x1(up[:, k].contiguous(), stage.weight[k], stage.bias[k]) for k in range(3)], dim=1)
    two = ctx.vote_frames(depth, maps[:, :, 1].contiguous(), K, max_radii=mr, mask_flags=api.RCV_MASK_MAX_RADIUS | api.RCV_MASK_RADIUS_POSITIVE)
    torch.cuda.synchronize()
    assert torch.equal(fused["radius"], maps[:, :, 1])
    for k in ("centre_mm", "peak", "votes", "n_points", "grid", "status"):
        assert torch.equal(fused[k], two[k]), k
    out = stage.vote(rgb, depth, K, mr, mask_flags=api.RCV_MASK_MAX_RADIUS | api.RCV_MASK_RADIUS_POSITIVE)
    assert torch.equal(out["n_points"], fused["n_points"])
    with torch.no_grad():
        up32 = trunks[0].float()(rgb.cuda()[:1])
    err = (up[:1, 0].float() - up32).abs().max().item()
    assert err <= 0.1 * max(1.0, up32.abs().max().item()), err


@_pending_gpu
def test_cuda_graphed_stage_equals_eager():
    """The CUDA-graphed producer (three trunks on forked streams inside one graph) gives the eager activations, replay after replay."""
    from rcvpose_b200 import api
    ctx = api.VoteContext(0, max_items=8, max_points_total=1 << 20, max_grid=256)
    torch.manual_seed(7)
    stage = producer.ProducerStage([producer.RadiusTrunk() for _ in range(3)], ctx)
    rgb = [torch.rand((1, 3, 96, 128)) for _ in range(3)]
    want = [stage.activations(x).clone() for x in rgb]
    stage.capture(1, 96, 128, concurrent=True)
    for x, w in zip(rgb + rgb[:1], want + want[:1]):
        got = stage.activations(x)
        torch.cuda.synchronize()
        assert got.shape == w.shape and torch.allclose(got.float(), w.float(), rtol=2e-2, atol=2e-2)     # cuDNN may pick another algorithm under capture
    assert stage.activations(torch.rand((2, 3, 96, 128))).shape[0] == 2                                      # another shape runs eagerly


@_pending_gpu
def test_lm_evaluator_checkpoint_branch_through_the_fused_stage(tmp_path):
    """N1 uses N2: evaluate_lm_class(using_ckpts=True, producer=<ProducerStage>) -- RGB -> trunks -> fused conv8 + mask rule + vote,
    ICP scene from the survival bits -- against the same evaluator fed sem / radius MAPS that rcv_head_1x1 produced from the same
    activations (the two-call route): keypoints, poses, ADD, scene sizes and ICP poses identical."""
    from PIL import Image
    from rcvpose_b200 import api, evaluate, formats, synth
    root = str(tmp_path) + "/"
    cls = "cat"
    stems = synth.write_lm_dataset(root, cls, 3, seed=5)
    rng = np.random.default_rng(0)
    for s in os.listdir(root + "LINEMOD/" + cls + "/JPEGImages/"):
        Image.fromarray(rng.integers(0, 256, size=(480, 640, 3)).astype(np.uint8)).save(root + "LINEMOD/" + cls + "/JPEGImages/" + s, quality=95)
    ctx = api.VoteContext(0, max_items=16, max_points_total=1 << 22, max_grid=400, image=(480, 640), max_model_points=1500)
    torch.manual_seed(11)
    trunks = [producer.RadiusTrunk() for _ in range(3)]
    for k, t in enumerate(trunks):          # untrained networks: make the head say "object everywhere, radius 1.1 .. 1.3 dm" so that the mask rule has work
        with torch.no_grad():
            t.conv8.weight.zero_()
            t.conv8.bias.copy_(torch.tensor([3.0, 1.1 + 0.1 * k]))
    stage = producer.ProducerStage(trunks, ctx)

    def maps(class_name, k, image_path):      # the two-call route: the same activations through rcv_head_1x1, maps on the host
        img = producer.normalise_rgb(formats.read_rgb(image_path))
        up = stage.activations(img)
        m = ctx.head_1x1(up[:, k - 1].contiguous(), stage.weight[k - 1], stage.bias[k - 1])
        return m[0, 0].cpu().numpy(), m[0, 1].cpu().numpy()

    a = evaluate.evaluate_lm_class(root, cls, using_ckpts=True, producer=stage, frames_per_batch=2, verbose=False)
    b = evaluate.evaluate_lm_class(root, cls, using_ckpts=True, producer=maps, frames_per_batch=2, verbose=False)
    assert a["frames"] == b["frames"] == sorted(stems)
    for key in ("centre_mm", "n_points", "peak", "scene_points", "icp_iters"):
        assert np.array_equal(a[key], b[key]), key
    for key in ("RT", "dist_before", "RT_icp", "dist_after"):
        np.testing.assert_allclose(a[key], b[key], rtol=1e-12, atol=1e-9, err_msg=key)
    assert (a["n_points"] > 1000).all()


def _tail_reference(x_bf16, w7, scale, shift, w8, b8):
    """conv7 + folded BN + ReLU + conv8 in float64 PyTorch on the operands the kernel sees (x, w7, w8 rounded to bf16; the
    activation rounded to bf16 before conv8): models/fcnresnet.py:114-118, :183-189."""
    x = x_bf16.double()
    a = F.conv2d(x, w7.bfloat16().double(), padding=1) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    a = torch.relu(a).float().bfloat16().double()
    return F.conv2d(a, w8.bfloat16().double().view(2, 32, 1, 1)) + b8.double().view(1, 2, 1, 1)


@_pending_gpu
@pytest.mark.parametrize("shape", [(1, 8, 128), (2, 19, 256), (1, 480, 640)])
def test_conv7_head_kernel_vs_torch(shape):
    """K6 (conv7head.cu): the implicit-GEMM tail against PyTorch float64 on the same bf16 operands, image borders (zero padding)
    included.  fp32 accumulation order can flip the bf16 rounding of an activation (2^-8 relative on one of 32 terms)."""
    from rcvpose_b200 import api
    B, H, W = shape
    ctx = api.VoteContext(0, max_items=8, max_points_total=1 << 20, max_grid=256)
    g = torch.Generator(device="cpu").manual_seed(11 + H)
    x = (torch.randn((B, 64, H, W), generator=g) * 0.7).bfloat16().cuda()
    w7 = (torch.randn((32, 64, 3, 3), generator=g) * 0.06).cuda()
    scale = (0.5 + torch.rand((32,), generator=g)).cuda()
    shift = (torch.randn((32,), generator=g) * 0.3).cuda()
    w8 = (torch.randn((2, 32), generator=g) * 0.3).cuda()
    b8 = torch.randn((2,), generator=g).cuda()
    got = ctx.conv7_head(x, w7, scale, shift, w8, b8)
    want = _tail_reference(x, w7, scale, shift, w8, b8)
    torch.cuda.synchronize()
    assert got.shape == (B, 2, H, W) and got.dtype == torch.float32
    err = (got.double() - want).abs()
    tol = 2.0 ** -8 * 4.0 * float(want.abs().max())
    assert float(err.max()) <= tol, (float(err.max()), tol)
    assert float(err.mean()) <= 1e-4 * float(want.abs().mean() + 1.0), float(err.mean())
    # channels_last input is used in place, an NCHW tensor is converted: same result
    got2 = ctx.conv7_head(x.contiguous(memory_format=torch.channels_last), w7, scale, shift, w8, b8)
    assert torch.equal(got, got2)


@_pending_gpu
def test_conv7_head_vote_frames_equals_the_two_call_route():
    """rcv_conv7_head_vote_frames (tail + mask rule in the epilogue) against rcv_conv7_head + rcv_vote_frames on the maps it wrote:
    same kernel arithmetic for seg / radial, so radius planes, survivors, centres and peaks must be identical."""
    from rcvpose_b200 import api, synth
    ctx = api.VoteContext(0, max_items=16, max_points_total=1 << 22, max_grid=400)
    frames = [synth.config3_frame(f) for f in (51, 52)]
    depth = torch.from_numpy(np.stack([f["depth"] for f in frames]).view(np.int16)).cuda()
    K = torch.from_numpy(frames[0]["K"]).cuda()
    g = torch.Generator(device="cpu").manual_seed(3)
    Kp = 3
    xs = [(torch.randn((2, 64, 480, 640), generator=g) * 0.5).bfloat16().cuda().contiguous(memory_format=torch.channels_last) for _ in range(Kp)]
    w7 = (torch.randn((Kp, 32, 64, 3, 3), generator=g) * 0.05).cuda()
    scale = (0.5 + torch.rand((Kp, 32), generator=g)).cuda()
    shift = (torch.randn((Kp, 32), generator=g) * 0.2).cuda()
    w8 = (torch.randn((Kp, 2, 32), generator=g) * 0.2).cuda()
    b8 = torch.tensor([[0.5, 1.2], [0.4, 1.0], [0.6, 1.4]]).cuda()      # seg around 0.5 (threshold), radius around 1.2 dm
    mr = torch.full((Kp,), 2.0, dtype=torch.float64, device="cuda")
    flags = api.RCV_MASK_MAX_RADIUS | api.RCV_MASK_SEM_GT | api.RCV_MASK_RADIUS_POSITIVE
    fused = ctx.conv7_head_vote_frames(xs, w7, scale, shift, w8, b8, depth, K, max_radii=mr, mask_flags=flags, sem_threshold=0.5, want_radius=True)
    maps = torch.stack([ctx.conv7_head(xs[k], w7[k], scale[k], shift[k], w8[k], b8[k]) for k in range(Kp)], dim=1)      # (B,Kp,2,H,W)
    two = ctx.vote_frames(depth, maps[:, :, 1].contiguous(), K, sem=maps[:, :, 0].contiguous(), max_radii=mr, mask_flags=flags, sem_threshold=0.5)
    torch.cuda.synchronize()
    assert torch.equal(fused["radius"], maps[:, :, 1])
    assert int(fused["n_points"].min()) > 100, fused["n_points"]
    for k in ("centre_mm", "peak", "votes", "n_points", "grid", "status"):
        assert torch.equal(fused[k], two[k]), k


@_pending_gpu
def test_producer_stage_with_the_fused_tail():
    """ProducerStage(fuse_tail=True): trunks up to up1, then ONE kernel for conv7 + BN + ReLU + conv8 + mask rule -- against the
    PyTorch float32 forward of the same trunks (forward_reference) to bf16 accuracy, and through the vote."""
    from rcvpose_b200 import api, synth
    ctx = api.VoteContext(0, max_items=16, max_points_total=1 << 22, max_grid=400)
    torch.manual_seed(9)
    trunks = [producer.RadiusTrunk() for _ in range(3)]
    with torch.no_grad():
        for t in trunks:                                  # non-trivial BatchNorm statistics in conv7
            t.conv7[1].running_mean.normal_(0.0, 0.1); t.conv7[1].running_var.uniform_(0.5, 1.5); t.conv7[1].weight.uniform_(0.5, 1.5); t.conv7[1].bias.normal_(0.0, 0.1)
    rgb = torch.rand((1, 3, 96, 128))
    with torch.no_grad():
        want = [torch.cat(t.float().cuda().eval().forward_reference(rgb.cuda()), 1) for t in trunks]      # (1,2,H,W) float32 each
    stage = producer.ProducerStage(trunks, ctx, fuse_tail=True)
    xs = stage.activations(rgb)
    assert len(xs) == 3 and xs[0].shape == (1, 64, 96, 128) and xs[0].is_contiguous(memory_format=torch.channels_last)
    w7, sc, sh, w8, b8 = stage.tail_params
    for k in range(3):
        got = ctx.conv7_head(xs[k], w7[k], sc[k], sh[k], w8[k], b8[k])
        torch.cuda.synchronize()
        err = (got - want[k]).abs().max().item()
        assert err <= 0.1 * max(1.0, want[k].abs().max().item()), (k, err)
    frames = [synth.config3_frame(61)]
    depth = torch.from_numpy(np.stack([f["depth"] for f in frames]).view(np.int16)).cuda()
    K = torch.from_numpy(frames[0]["K"]).cuda()
    mr = torch.full((3,), 1.0e9, dtype=torch.float64, device="cuda")
    rgb2 = torch.rand((1, 3, 480, 640))
    out = stage.vote(rgb2, depth, K, mr, mask_flags=api.RCV_MASK_MAX_RADIUS)
    torch.cuda.synchronize()
    assert out["centre_mm"].shape == (1, 3, 3) and bool((out["status"] >= 0).all())
    stage.capture(1, 480, 640)
    out2 = stage.vote(rgb2, depth, K, mr, mask_flags=api.RCV_MASK_MAX_RADIUS)
    torch.cuda.synchronize()
    assert torch.equal(out2["status"], out["status"])
