"""MobileViT-xs layer graph (reference models/mobile_vit.py): unquantized 3x3/s2 stem conv, MV2
inverted-residual blocks (quantized 1x1 expand -> quantized depthwise 3x3 -> quantized 1x1 project,
SiLU), three MobileViT blocks (quantized 3x3 and 1x1 convs around an UNquantized nn.Linear
transformer over patch groups, then a quantized 1x1 and a quantized 3x3 fusion conv on the
concatenation), unquantized 1x1 head conv, mean pool, bias-free linear classifier.
BASELINE.json configs[3]: 224x224 with patch_size (1,1) built directly (the reference's factory
cannot build a working 224x224 model), or 256x256 with patch_size (2,2)."""
import torch
import torch.nn as nn

CHANNELS = (16, 32, 48, 48, 64, 64, 80, 80, 96, 96, 384)
DIMS = (96, 120, 144)
DEPTHS = (2, 4, 3)


_FUSED_NORM = [False]       # set by the factory while it builds a model


def _norm(c, act):
    """[norm (, SiLU)]; the fused norm applies the SiLU itself (forward-only path) and an nn.Identity keeps
    the reference's nn.Sequential indices, i.e. the state_dict keys"""
    if _FUSED_NORM[0]:
        from po2_quantization_b200 import FusedSyncBatchNorm
        return [FusedSyncBatchNorm(c, act="silu" if act else None)] + ([nn.Identity()] if act else [])
    return [nn.SyncBatchNorm(c)] + ([nn.SiLU()] if act else [])


def _cbs(conv, c_out):
    return nn.Sequential(conv, *_norm(c_out, True))


class _FeedForward(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden), nn.SiLU(), nn.Dropout(0.0),
                                 nn.Linear(hidden, dim), nn.Dropout(0.0))

    def forward(self, x):
        return self.net(x)


class _Attention(nn.Module):
    def __init__(self, dim, heads=4, dim_head=8):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(0.0)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(0.0))

    def forward(self, x):                                     # x: (b, p, n, dim)
        b, p, n, _ = x.shape
        q, k, v = self.to_qkv(self.norm(x)).chunk(3, dim=-1)
        split = lambda t: t.reshape(b, p, n, self.heads, -1).permute(0, 1, 3, 2, 4)    # b p h n d
        q, k, v = split(q), split(k), split(v)
        attn = self.attend(torch.matmul(q, k.transpose(-1, -2)) * self.scale)
        out = torch.matmul(attn, v).permute(0, 1, 3, 2, 4).reshape(b, p, n, -1)
        return self.to_out(out)


class _Transformer(nn.Module):
    def __init__(self, dim, depth, mlp_dim):
        super().__init__()
        self.layers = nn.ModuleList([nn.ModuleList([_Attention(dim), _FeedForward(dim, mlp_dim)]) for _ in range(depth)])

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return x


class _MV2(nn.Module):
    def __init__(self, conv_cls, c_in, c_out, stride, expansion, q):
        super().__init__()
        hid = int(c_in * expansion)
        self.use_res_connect = stride == 1 and c_in == c_out
        layers = []
        if expansion != 1:
            layers += [conv_cls(c_in, hid, 1, 1, 0, bias=False, **q)] + _norm(hid, True)
        layers += [conv_cls(hid, hid, 3, stride, 1, groups=hid, bias=False, **q)] + _norm(hid, True)
        layers += [conv_cls(hid, c_out, 1, 1, 0, bias=False, **q)] + _norm(c_out, False)
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        out = self.conv(x)
        return out + x if self.use_res_connect else out


class _ViTBlock(nn.Module):
    def __init__(self, conv_cls, dim, depth, channel, patch, mlp_dim, q):
        super().__init__()
        self.ph, self.pw = patch
        self.conv1 = _cbs(conv_cls(channel, channel, 3, 1, 1, bias=False, **q), channel)
        self.conv2 = _cbs(conv_cls(channel, dim, 1, 1, 0, bias=False, **q), dim)
        self.transformer = _Transformer(dim, depth, mlp_dim)
        self.conv3 = _cbs(conv_cls(dim, channel, 1, 1, 0, bias=False, **q), channel)
        self.conv4 = _cbs(conv_cls(2 * channel, channel, 3, 1, 1, bias=False, **q), channel)

    def forward(self, x):
        y = x.clone()
        x = self.conv2(self.conv1(x))
        b, d, H, W = x.shape
        h, w, ph, pw = H // self.ph, W // self.pw, self.ph, self.pw
        # "b d (h ph) (w pw) -> b (ph pw) (h w) d"
        x = x.reshape(b, d, h, ph, w, pw).permute(0, 3, 5, 2, 4, 1).reshape(b, ph * pw, h * w, d)
        x = self.transformer(x)
        # "b (ph pw) (h w) d -> b d (h ph) (w pw)"
        x = x.reshape(b, ph, pw, h, w, d).permute(0, 5, 3, 1, 4, 2).reshape(b, d, H, W)
        x = self.conv3(x)
        return self.conv4(torch.cat((x, y), 1))


class _MeanHW(nn.Module):
    def forward(self, x):
        return x.mean(dim=(2, 3))


class MobileViTXS(nn.Module):
    def __init__(self, conv_cls, image_size, num_classes, patch_size, quantize_fn, bits, expansion=4):
        super().__init__()
        ch, q = CHANNELS, dict(quantize_fn=quantize_fn, bits=bits)
        assert image_size[0] % patch_size[0] == 0 and image_size[1] % patch_size[1] == 0
        self.conv1 = _cbs(nn.Conv2d(3, ch[0], 3, 2, 1, bias=False), ch[0])            # never quantized
        self.stem = nn.ModuleList([_MV2(conv_cls, ch[0], ch[1], 1, expansion, q), _MV2(conv_cls, ch[1], ch[2], 2, expansion, q),
                                   _MV2(conv_cls, ch[2], ch[3], 1, expansion, q), _MV2(conv_cls, ch[2], ch[3], 1, expansion, q)])
        self.trunk = nn.ModuleList([
            nn.ModuleList([_MV2(conv_cls, ch[3], ch[4], 2, expansion, q), _ViTBlock(conv_cls, DIMS[0], DEPTHS[0], ch[5], patch_size, int(DIMS[0] * 2), q)]),
            nn.ModuleList([_MV2(conv_cls, ch[5], ch[6], 2, expansion, q), _ViTBlock(conv_cls, DIMS[1], DEPTHS[1], ch[7], patch_size, int(DIMS[1] * 4), q)]),
            nn.ModuleList([_MV2(conv_cls, ch[7], ch[8], 2, expansion, q), _ViTBlock(conv_cls, DIMS[2], DEPTHS[2], ch[9], patch_size, int(DIMS[2] * 4), q)]),
        ])
        self.to_logits = nn.Sequential(_cbs(nn.Conv2d(ch[-2], ch[-1], 1, 1, 0, bias=False), ch[-1]), _MeanHW(),
                                       nn.Linear(ch[-1], num_classes, bias=False))

    def forward(self, x):
        x = self.conv1(x)
        for blk in self.stem:
            x = blk(x)
        for mv2, vit in self.trunk:
            x = vit(mv2(x))
        return self.to_logits(x)


def mobilevit_xs(image_size=(224, 224), num_classes=1000, patch_size=(1, 1), quantize_fn=None, bits=8, conv_cls=None,
                 fused_norm=None):
    """fused_norm: FusedSyncBatchNorm (norm + SiLU in one kernel when no gradient is needed) instead of
    nn.SyncBatchNorm + nn.SiLU; default: on with this repo's conv class, off when another class is passed"""
    if fused_norm is None:
        fused_norm = conv_cls is None
    if conv_cls is None:
        from po2_quantization_b200 import QuantizedConv2d as conv_cls
    _FUSED_NORM[0] = bool(fused_norm)
    try:
        return MobileViTXS(conv_cls, tuple(image_size), num_classes, tuple(patch_size), quantize_fn, bits)
    finally:
        _FUSED_NORM[0] = False
