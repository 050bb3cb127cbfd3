"""``UNetModified2`` — host mirror of reference model/UNetModified2.py:146-269.

The module is a *parameter container*: it creates the same torch layers, in the same order and under the same
attribute names as the reference, so (a) ``state_dict()`` keys / shapes are identical and reference checkpoints load
unchanged, and (b) default initialisation under a given ``torch.manual_seed`` yields the same weights.  It contains
no torch compute: ``forward`` hands the raw pointers to the CUDA plan (csrc/, through include/sddm_b200.h), which
runs framing + concat + the whole UNet + overlap-add.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from ..plan import Plan, default_precision


class _Holder(nn.Module):
    """Parameter-only module: calling it is a bug (all compute lives in the CUDA plan)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sddm_b200 host modules hold parameters only; compute runs in the CUDA plan")


class PositionalEncoding(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        step = torch.arange(dim // 2)
        self.embedding_vector = 1e4 * 10.0 ** (-step * 4.0 / (dim // 2))   # plain attribute, as in the reference (:55)


class Swish(_Holder):
    pass


class FeatureWiseAffine(_Holder):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.noise_func = nn.Sequential(nn.Linear(in_channels, out_channels))


class Block(_Holder):
    def __init__(self, dim, dim_out, groups):
        super().__init__()
        self.block = nn.Sequential(nn.GroupNorm(groups, dim), Swish(), nn.Identity(), nn.Conv2d(dim, dim_out, 3, padding=1))


class ResnetBlock(_Holder):
    def __init__(self, dim, dim_out, noise_level_emb_dim, norm_groups):
        super().__init__()
        self.noise_func = FeatureWiseAffine(noise_level_emb_dim, dim_out)
        self.block1 = Block(dim, dim_out, norm_groups)
        self.block2 = Block(dim_out, dim_out, norm_groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


class Downsample(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, 3, 2, 1)


class Upsample(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, 3, padding=1)


class UNetModified2(nn.Module):
    def __init__(self, num_samples, in_channel=2, out_channel=1, inner_channel=32, norm_groups=32,
                 channel_mults=(1, 2, 3, 4, 5), res_blocks=3, dropout=0, segment_len=128, segment_stride=64):
        super().__init__()
        assert (num_samples - segment_len) % segment_stride == 0        # reference UNetModified2.py:13
        if dropout != 0:
            raise NotImplementedError("dropout != 0 is a training-only option; the inference hot path uses 0")
        self.cfg = dict(num_samples=num_samples, in_channel=in_channel, out_channel=out_channel,
                        inner_channel=inner_channel, norm_groups=norm_groups, channel_mults=tuple(channel_mults),
                        res_blocks=res_blocks, dropout=dropout, segment_len=segment_len, segment_stride=segment_stride)
        emb = inner_channel
        self.noise_level_mlp = nn.Sequential(PositionalEncoding(emb), nn.Linear(emb, emb * 4), Swish(),
                                             nn.Linear(emb * 4, emb), Swish())
        self.downs = nn.ModuleList([nn.Conv2d(in_channel, inner_channel, kernel_size=3, padding=1)])
        feat, cin = [inner_channel], inner_channel
        for mult in channel_mults:
            cout = inner_channel * mult
            for _ in range(res_blocks):
                self.downs.append(ResnetBlock(cin, cout, emb, norm_groups))
                feat.append(cout)
                cin = cout
            self.downs.append(Downsample(cout))
            feat.append(cout)
        self.mid = nn.ModuleList([ResnetBlock(cin, cin, emb, norm_groups)])
        self.ups = nn.ModuleList([])
        cout = cin
        for lvl in reversed(range(len(channel_mults))):
            cin = inner_channel * channel_mults[lvl]
            cout = cin
            self.ups.append(ResnetBlock(cin + feat.pop(), cout, emb, norm_groups))
            self.ups.append(Upsample(cout))
            cout = inner_channel if lvl == 0 else inner_channel * channel_mults[lvl - 1]
            for _ in range(res_blocks):
                self.ups.append(ResnetBlock(cin + feat.pop(), cout, emb, norm_groups))
                cin = cout
        self.final_conv = Block(cout, out_channel, norm_groups)
        self.precision: Optional[int] = None     # None -> SDDM_B200_PRECISION env / default
        self._plans: Dict[tuple, Plan] = {}

    # -- plan management -------------------------------------------------------------------------------
    def _param_version(self):
        return tuple(int(p._version) for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())

    def invalidate_plans(self):
        self._plans = {}

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plans = {}
        return super().load_state_dict(*a, **k)

    def get_plan(self, diffusion=None, precision: Optional[int] = None) -> Plan:
        """Plan for (current weights, diffusion schedule, precision) on the weights' device; cached."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("UNetModified2 (sddm_b200) must live on a CUDA device: call .to('cuda') first; "
                               "there is no CPU fallback")
        prec = precision if precision is not None else (self.precision if self.precision is not None else default_precision())
        tables = diffusion.host_tables() if diffusion is not None else None
        key = (diffusion.tables_key() if diffusion is not None else None, prec, str(dev), self._param_version())
        plan = self._plans.get(key)
        if plan is None:
            self._plans = {k: v for k, v in self._plans.items() if k[3] == key[3]}   # drop plans of stale weights
            if diffusion is not None:
                T = diffusion.num_timesteps
            else:   # forward() alone takes explicit noise levels; the per-t table is unused
                import numpy as np
                from .._lib import SCHEDULE_FIELDS
                T = 1
                tables = {k: np.ones(2, dtype=np.float32) for k in SCHEDULE_FIELDS}
            weights = {k: v for k, v in self.state_dict().items()}
            plan = Plan(self.cfg, weights, tables, T, prec, dev, pe_vector=self.noise_level_mlp[0].embedding_vector)
            self._plans[key] = plan
        return plan

    # -- reference API ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, y_t, diffusion_step):
        """eps_hat[B,1,T] from condition x[B,1,T], iterate y_t[B,1,T], noise level [B,1,1] (reference :237-269)."""
        return self.get_plan().eps(x, y_t, noise_level=diffusion_step).reshape(x.shape)
