"""Self-contained CPU restatement of the reference models (TEST INFRASTRUCTURE; travels to the GPU box).

Module / parameter names equal the reference's so a reference-layout state_dict loads directly.  Composition follows
the reference files line by line; third-party blocks come from ref_timm.py / ref_monai.py.  Pinned in this container
against the reference's own code (oracle/reference_loader.py) by tests/test_oracle.py and tests/golden/.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ref_monai, ref_timm


def icnr_init(conv, upsample_factor, upsample_dims, init=nn.init.kaiming_normal_):
    # VM/components/blocks.py:14-51
    out_channels, in_channels, *dims = conv.weight.shape
    scale_factor = upsample_factor**upsample_dims
    oc2 = int(out_channels / scale_factor)
    kernel = init(torch.zeros([oc2, in_channels] + dims)).transpose(0, 1)
    kernel = kernel.reshape(oc2, in_channels, -1).repeat(1, 1, scale_factor)
    kernel = kernel.reshape([in_channels, out_channels] + dims).transpose(0, 1)
    conv.weight.data.copy_(kernel)


class Stem(nn.Module):
    # VM/components/stems.py:26-50 (UNeXt2Stem) and :69-74,117-134 (StemDepthtoChannels): Conv3d then fold D into C
    def __init__(self, in_channels, out_channels_3d, kernel_size, stride):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels_3d, kernel_size=kernel_size, stride=stride)

    def forward(self, x):
        x = self.conv(x)
        b, c, d, h, w = x.shape
        return x.reshape(b, c * d, h, w)


class UpStage(nn.Module):
    # VM/components/blocks.py:136-172 (pixelshuffle branch)
    def __init__(self, in_channels, skip_channels, out_channels, scale_factor, conv_blocks):
        super().__init__()
        mid = in_channels // scale_factor**2
        self.upsample = ref_monai.UpSample(2, in_channels, mid, scale_factor, "pixelshuffle", None, False)
        # VM/components/blocks.py:54-74
        self.conv = ref_timm.ConvNeXtStage(mid + skip_channels, out_channels, stride=1, depth=conv_blocks,
                                           ls_init_value=None, conv_mlp=True, use_grn=True,
                                           norm_layer=ref_timm.LayerNorm2d, norm_layer_cl=ref_timm.LayerNorm)
        self.conv.apply(ref_timm._init_weights)
        icnr_init(self.conv.blocks[-1].mlp.fc2, scale_factor, 2)

    def forward(self, inp, skip):
        return self.conv(torch.cat([self.upsample(inp), skip], dim=1))


class Decoder(nn.Module):
    # VM/components/blocks.py:197-243
    def __init__(self, num_channels, conv_blocks, strides):
        super().__init__()
        self.decoder_stages = nn.ModuleList(
            [UpStage(num_channels[i], num_channels[i] // 2, num_channels[i + 1], strides[i], conv_blocks)
             for i in range(len(num_channels) - 1)])

    def forward(self, features):
        feat = features[0]
        for skip, stage in zip(list(features[1:]) + [None], self.decoder_stages):
            feat = stage(feat, skip)
        return feat


class Head(nn.Module):
    # VM/components/heads.py:597-641
    def __init__(self, in_channels, out_channels, out_stack_depth, expansion_ratio, pool):
        super().__init__()
        self.upsample = ref_monai.UpSample(2, in_channels, in_channels // 4, 2, "pixelshuffle", None, pool)
        mid = out_channels * expansion_ratio * 4
        self.conv = nn.Sequential(
            ref_monai.Convolution(3, in_channels // 4 // (out_stack_depth + 2), mid, kernel_size=3, padding=(0, 1, 1)),
            nn.Conv3d(mid, out_channels * 4, 1))
        icnr_init(self.conv[-1], 2, 2)
        self.out = nn.PixelShuffle(2)
        self.out_stack_depth = out_stack_depth

    def forward(self, x):
        x = self.upsample(x)
        d = self.out_stack_depth + 2
        b, c, h, w = x.shape
        x = self.conv(x.reshape((b, c // d, d, h, w)))
        return self.out(x.transpose(1, 2)).transpose(1, 2)


class UNeXt2(nn.Module):
    # VM/unet/unext2.py:16-82
    def __init__(self, in_channels=1, out_channels=1, in_stack_depth=5, out_stack_depth=None,
                 backbone="convnextv2_tiny", stem_kernel_size=(5, 4, 4), decoder_conv_blocks=2, head_pool=False,
                 head_expansion_ratio=4, drop_path_rate=0.0, **ignored):
        super().__init__()
        if out_stack_depth is None:
            out_stack_depth = in_stack_depth
        enc = ref_timm.create_model(backbone, features_only=True, drop_path_rate=drop_path_rate)
        ch = enc.feature_info.channels()
        enc.stem_0 = nn.Identity()
        self.encoder_stages = enc
        ratio = in_stack_depth // stem_kernel_size[0]
        self.stem = Stem(in_channels, ch[0] // ratio, stem_kernel_size, stem_kernel_size)
        ch.reverse()
        ch[-1] = (out_stack_depth + 2) * out_channels * 4 * head_expansion_ratio
        self.decoder = Decoder(ch, decoder_conv_blocks, [2] * (len(ch) - 1) + [stem_kernel_size[-1]])
        self.head = Head(ch[-1], out_channels, out_stack_depth, head_expansion_ratio, head_pool)

    def forward(self, x):
        x = self.encoder_stages(self.stem(x))
        x.reverse()
        return self.head(self.decoder(x))


class ContrastiveEncoder(nn.Module):
    # VM/contrastive/encoder.py:79-154 (convnext backbones)
    def __init__(self, backbone, in_channels, in_stack_depth, stem_kernel_size=(5, 4, 4), stem_stride=(5, 4, 4),
                 embedding_dim=768, projection_dim=128, drop_path_rate=0.0, **ignored):
        super().__init__()
        enc = ref_timm.create_model(backbone, features_only=False, drop_path_rate=drop_path_rate,
                                    num_classes=embedding_dim)
        c_enc = enc.stem[0].out_channels
        enc.stem[0] = nn.Identity()
        self.projection_src = None
        projection = nn.Sequential(nn.Linear(enc.num_features, embedding_dim), nn.BatchNorm1d(embedding_dim),
                                   nn.ReLU(inplace=True), nn.Linear(embedding_dim, projection_dim),
                                   nn.BatchNorm1d(projection_dim))
        enc.head.fc = nn.Identity()
        d_out = (in_stack_depth - stem_kernel_size[0]) // stem_stride[0] + 1
        self.stem = Stem(in_channels, c_enc // d_out, stem_kernel_size, stem_stride)
        self.encoder = enc
        self.projection = projection
        del self.projection_src

    def forward(self, x):
        emb = self.encoder(self.stem(x))
        return emb, self.projection(emb)


def ntxent(embeddings, labels, temperature=0.07):
    """pytorch-metric-learning 2.9 NTXentLoss with CosineSimilarity + MeanReducer (SURVEY Appendix B.4;
    in-tree spec VM/contrastive/loss.py:135-185 with hcl weights == 1)."""
    e = torch.nn.functional.normalize(embeddings.float(), dim=1)
    sim = e @ e.t()
    same = labels[:, None] == labels[None, :]
    eye = torch.eye(len(labels), dtype=torch.bool, device=sim.device)
    a1, p = torch.where(same & ~eye)
    a2, n = torch.where(~same)
    pos = sim[a1, p].unsqueeze(1) / temperature
    neg = sim[a2, n] / temperature
    n_per_p = (a2.unsqueeze(0) == a1.unsqueeze(1)).float()
    neg_m = neg * n_per_p
    neg_m[n_per_p == 0] = torch.finfo(neg.dtype).min
    mx = torch.max(pos, neg_m.max(dim=1, keepdim=True)[0]).detach()
    num = torch.exp(pos - mx).squeeze(1)
    den = torch.sum(torch.exp(neg_m - mx), dim=1) + num
    return (-torch.log(num / den + torch.finfo(neg.dtype).tiny)).mean()
