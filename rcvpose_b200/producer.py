"""The radius-map producer as the INPUT STAGE of the voting path (BASELINE.json configs[1]).

The north star keeps the FCN-ResNet of the reference (`DenseFCNResNet152`, models/fcnresnet.py:47-191) in PyTorch bf16; only its
last layer, the 1x1 `conv8`, is a CUDA kernel of this package (rcv_head_1x1 / rcv_head_vote_frames).  This module is that
PyTorch stage: `RadiusTrunk` is the network up to and including conv7 + BN + ReLU (the 32-channel map the head consumes), with
the reference's parameter names so that a reference checkpoint loads with `load_state_dict` unchanged; conv8's weight and bias
are exposed as plain tensors for the head kernel.  `AccumulatorSpace.FCResBackbone`'s preprocessing (:140-150) is
`normalise_rgb`.  Nothing here votes; nothing here is needed by the voting entry points.

Architecture (restated from the reference's layer list, not copied): ResNet-152-style encoder -- 7x7/2 stem, 3x3/2 max pool, four
stages of (1 + n) bottlenecks with n = 2, 7, 35, 2 and widths 64, 128, 256, 512 (x4 out), every bottleneck owning a projection
branch that only the first of a stage uses, 3x3 convolutions WITH bias -- a 3x3 2048 -> 1024 bridge, and a decoder of five
[concat skip -> 3x3 conv -> BN -> ReLU -> 2x bilinear up] steps down to 64 channels at full resolution, then conv7 (3x3, 64 -> 32).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_STAGES = ((64, 2, 1), (128, 7, 2), (256, 35, 2), (512, 2, 2))      # (width, extra bottlenecks, stride of the first)
_DECODER = ((5, 2048 + 1024, 1024), (4, 1024 + 1024, 512), (3, 512 + 512, 256), (2, 256 + 256, 128), (1, 64 + 128, 64))
_MEAN, _STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _conv_bn(cin, cout, k, stride=1, bias=True):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=bias), nn.BatchNorm2d(cout)


class _Bottleneck(nn.Module):
    """1x1 -> 3x3 (stride) -> 1x1 (x4) with a residual; `project` selects the 1x1 projection of the input as the residual."""

    def __init__(self, cin, width, stride, project):
        super().__init__()
        self.conv1, self.bn1 = _conv_bn(cin, width, 1, bias=False)
        self.conv2, self.bn2 = _conv_bn(width, width, 3, stride=stride, bias=True)
        self.conv3, self.bn3 = _conv_bn(width, width * 4, 1, bias=False)
        self.upsample_ = nn.Sequential(*_conv_bn(cin, width * 4, 1, stride=stride, bias=False))   # exists in every block of a checkpoint
        self.project = project

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + (self.upsample_(x) if self.project else x))


class RadiusTrunk(nn.Module):
    """DenseFCNResNet152 of the reference up to conv7 (B,3,H,W) -> (B,32,H,W); H, W multiples of 32.  `head()` returns conv8's
    (weight (2,32), bias (2,)) for rcv_head_1x1; `forward_reference` adds conv8 in PyTorch (parity tests only)."""

    def __init__(self):
        super().__init__()
        self.conv1, self.bn1 = _conv_bn(3, 64, 7, stride=2, bias=False)
        cin = 64
        for s, (width, extra, stride) in enumerate(_STAGES, start=1):
            setattr(self, "block%dup" % s, _Bottleneck(cin, width, stride, True))
            cin = width * 4
            setattr(self, "block%d" % s, nn.Sequential(*[_Bottleneck(cin, width, 1, False) for _ in range(extra)]))
        self.conv6, self.bn6 = _conv_bn(2048, 1024, 3)
        for level, c_in, c_out in _DECODER:
            setattr(self, "conv_up%d" % level, nn.Sequential(*_conv_bn(c_in, c_out, 3), nn.ReLU(inplace=True)))
        self.conv7 = nn.Sequential(*_conv_bn(64, 32, 3), nn.ReLU(inplace=True))
        self.conv8 = nn.Conv2d(32, 2, kernel_size=1)

    def forward(self, x):
        return self.conv7(self.forward_up1(x))

    def forward_up1(self, x):
        """(B,3,H,W) -> (B,64,H,W): everything before conv7 (the input of the fused tail kernel, rcv_conv7_head)."""
        stem = F.relu(self.bn1(self.conv1(x)))               # the reference's ReLU is in place, so its last skip is the ReLU'd stem (:121-123,180)
        skips = [stem]
        y = F.max_pool2d(stem, kernel_size=3, stride=2, padding=1)
        for s in range(1, 5):
            y = getattr(self, "block%d" % s)(getattr(self, "block%dup" % s)(y))
            skips.append(y)
        up = F.relu(self.bn6(self.conv6(y)))
        for level, _, _ in _DECODER:
            up = getattr(self, "conv_up%d" % level)(torch.cat((up, skips[level - 1]), 1))
            up = F.interpolate(up, scale_factor=2, mode="bilinear", align_corners=False)
        return up

    def head(self):
        return self.conv8.weight.detach().reshape(2, 32), self.conv8.bias.detach().reshape(2)

    def tail(self):
        """Parameters of the fused tail kernel (rcv_conv7_head): conv7's weight (32,64,3,3), its BatchNorm (eval mode) folded with
        the convolution's bias into scale / shift (32,), conv8's weight (2,32) and bias (2,) -- float32."""
        conv, bn = self.conv7[0], self.conv7[1]
        scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        cb = conv.bias.detach().double() if conv.bias is not None else torch.zeros_like(scale)
        shift = (cb - bn.running_mean.detach().double()) * scale + bn.bias.detach().double()
        w8, b8 = self.head()
        return conv.weight.detach().float(), scale.float(), shift.float(), w8.float(), b8.float()

    def forward_reference(self, x):
        out = self.conv8(self.forward(x))
        return out[:, :1], out[:, 1:]


def normalise_rgb(img_u8):
    """(H,W,3) or (B,H,W,3) uint8 RGB -> (B,3,H,W) float32, ImageNet mean / std (FCResBackbone, AccumulatorSpace.py:143-147)."""
    a = np.asarray(img_u8, dtype=np.float64)
    if a.ndim == 3:
        a = a[None]
    a = (a / 255.0 - np.array(_MEAN)) / np.array(_STD)
    return torch.from_numpy(np.ascontiguousarray(a.transpose(0, 3, 1, 2))).float()


class ProducerStage:
    """Three keypoint networks (AccumulatorSpace.py:516-527) in bf16 on one GPU, feeding the fused head + vote entry point:
    images -> conv7 activations (B,3,32,H,W) bf16 -> VoteContext.head_vote_frames.  The trunks are ordinary PyTorch modules in
    eval mode; the head and everything after it are this package's CUDA kernels.

    `capture(batch, H, W)` records the three trunk forwards of a fixed batch shape in ONE CUDA graph (the trunks on three
    forked streams, joined before the head): a 480x640 forward of one trunk is ~500 small cuDNN launches, launch-bound at the
    batch sizes an evaluator uses, and the reference runs them one image and one network at a time (:594-599)."""

    def __init__(self, trunks, ctx, dtype=torch.bfloat16, channels_last=False, fuse_tail=False):
        """fuse_tail: the trunks stop before conv7 and conv7 + BN + ReLU + conv8 + the mask rule run as ONE kernel of this package
        (rcv_conv7_head_vote_frames): the 32-channel activation never reaches HBM.  Needs W % 128 == 0 and uses channels_last."""
        assert len(trunks) >= 1
        self.ctx, self.dtype = ctx, dtype
        self.fuse_tail = bool(fuse_tail)
        channels_last = channels_last or self.fuse_tail
        self.trunks = [t.to(device=ctx.device, dtype=dtype).eval() for t in trunks]
        self.channels_last = bool(channels_last)
        if self.channels_last:
            self.trunks = [t.to(memory_format=torch.channels_last) for t in self.trunks]
        w, b = zip(*[t.head() for t in self.trunks])
        self.weight = torch.stack([x.float() for x in w]).contiguous()     # (Kp,2,32)
        self.bias = torch.stack([x.float() for x in b]).contiguous()       # (Kp,2)
        if self.fuse_tail:
            tails = [t.tail() for t in self.trunks]                         # from the trunk's parameters as the unfused modules see them (rounded to `dtype`), folded in float64
            self.tail_params = [torch.stack([tl[i].float() for tl in tails]).contiguous().to(ctx.device) for i in range(5)]
        self._graph = None

    @torch.no_grad()
    def _run(self, x, up, concurrent):
        """x (B,3,H,W) -> up (B,Kp,32,H,W) (NCHW planes per (frame, keypoint): what the head's TMA boxes read); with fuse_tail
        up is a list of Kp (B,64,H,W) channels_last tensors (the up1 outputs)."""
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        if self.fuse_tail:
            fwd = lambda k, t: up[k].copy_(t.forward_up1(x))  # noqa: E731
        else:
            def fwd(k, t):
                up[:, k] = t(x)
        if not concurrent or len(self.trunks) == 1:
            for k, t in enumerate(self.trunks):
                fwd(k, t)
            return
        main = torch.cuda.current_stream()
        side = self._side_streams
        for k, t in enumerate(self.trunks):
            st = main if k == 0 else side[k - 1]
            if k:
                st.wait_stream(main)
            with torch.cuda.stream(st):
                fwd(k, t)
        for st in side[:len(self.trunks) - 1]:
            main.wait_stream(st)

    @torch.no_grad()
    def capture(self, batch, H, W, concurrent=True):
        """CUDA-graphs the trunk forwards for images of shape (batch,3,H,W); afterwards `activations` replays the graph for
        inputs of that shape (and runs eagerly for any other)."""
        dev = self.ctx.device
        self._side_streams = [torch.cuda.Stream(device=dev) for _ in range(max(0, len(self.trunks) - 1))]
        self._static_in = torch.zeros((batch, 3, H, W), dtype=self.dtype, device=dev)
        self._static_up = self._alloc_up(batch, H, W)
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            for _ in range(3):                                   # cuDNN autotuning and workspace allocation happen outside the capture
                self._run(self._static_in, self._static_up, concurrent)
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._run(self._static_in, self._static_up, concurrent)
        self._graph = g
        return self

    def _alloc_up(self, B, H, W):
        dev = self.ctx.device
        if self.fuse_tail:
            return [torch.empty((B, 64, H, W), dtype=self.dtype, device=dev).contiguous(memory_format=torch.channels_last) for _ in self.trunks]
        return torch.empty((B, len(self.trunks), 32, H, W), dtype=self.dtype, device=dev)

    @torch.no_grad()
    def activations(self, images):
        """images (B,3,H,W) float -> (B,Kp,32,H,W) in the stage's dtype, NCHW-contiguous per (frame, keypoint)."""
        x = images.to(device=self.ctx.device, dtype=self.dtype)
        if self._graph is not None and tuple(x.shape) == tuple(self._static_in.shape):
            self._static_in.copy_(x)
            self._graph.replay()
            return self._static_up
        B, _, H, W = x.shape
        up = self._alloc_up(B, H, W)
        self._run(x, up, False)
        return up

    def vote(self, images, depth, K, max_radii, **kw):
        """RGB + depth -> keypoint centres: conv7 activations, then rcv_head_vote_frames (mask rule sem > 0.8, radial <= max_radii);
        with fuse_tail: up1 activations, then rcv_conv7_head_vote_frames."""
        if self.fuse_tail:
            w7, sc, sh, w8, b8 = self.tail_params
            return self.ctx.conv7_head_vote_frames(self.activations(images), w7, sc, sh, w8, b8, depth, K, max_radii=max_radii, **kw)
        return self.ctx.head_vote_frames(self.activations(images), self.weight, self.bias, depth, K, max_radii=max_radii, **kw)
