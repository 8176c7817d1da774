"""nn.Module surface of the reference (boundary b3) on top of the fused B200 operators.

  SS2D             <- /root/reference/gm-unet/model/gm/ss2d.py:521-556 (k_group=1, directions given per call) and
                      model/vmamba/vmamba.py:992 with forward_type "v2"/"v2_no32" (k_group=4)
  GroupMambaLayer  <- model/gm/groupmamba.py:85-159
  mamba_init       <- model/gm/ss2d.py:154-209

Parameter names, shapes, dtypes and the order in which the constructors draw random numbers are those of the
reference, so `load_state_dict` works in both directions. The forward differs only in WHERE the work happens:
cross-scan and cross-merge are addressing inside the scan kernels, out_norm + SiLU gate are one fused pass, and
no permuted or transposed copy is materialised.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn


class mamba_init:
    """Parameter initialisers (ss2d.py:154-209), same RNG consumption."""

    @staticmethod
    def dt_init(dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4):
        dt_proj = nn.Linear(dt_rank, d_inner, bias=True)
        std = dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(dt_proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(d_inner) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(min=dt_init_floor)
        inv_dt = dt + torch.log(-torch.expm1(-dt))            # softplus^-1
        with torch.no_grad():
            dt_proj.bias.copy_(inv_dt)
        return dt_proj

    @staticmethod
    def A_log_init(d_state, d_inner, copies=-1, device=None, merge=True):
        A_log = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)).repeat(d_inner, 1).contiguous()
        if copies > 0:
            A_log = A_log.unsqueeze(0).repeat(copies, 1, 1)
            if merge:
                A_log = A_log.flatten(0, 1)
        A_log = nn.Parameter(A_log)
        A_log._no_weight_decay = True
        return A_log

    @staticmethod
    def D_init(d_inner, copies=-1, device=None, merge=True):
        D = torch.ones(d_inner, device=device)
        if copies > 0:
            D = D.unsqueeze(0).repeat(copies, 1)
            if merge:
                D = D.flatten(0, 1)
        D = nn.Parameter(D)
        D._no_weight_decay = True
        return D


class SS2D(nn.Module, mamba_init):
    """2D selective-scan block: in_proj -> dwconv3x3 -> SiLU -> [K-direction selective scan] -> out_norm ->
    * SiLU(z) -> out_proj.  forward(x: (B, H, W, C), CrossScan=None, CrossMerge=None) -> (B, H, W, C).

    k_group=1 reproduces model/gm/ss2d.py (the direction comes from the (CrossScan_k, CrossMerge_k) classes passed
    to forward, as GroupMambaLayer does); k_group=4 reproduces the VMamba SS2D "v2" (four directions, no arguments).
    """

    def __init__(self, d_model=96, d_state=16, ssm_ratio=2.0, dt_rank="auto", act_layer=nn.SiLU, d_conv=3,
                 conv_bias=True, dropout=0.0, bias=False, dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0,
                 dt_init_floor=1e-4, initialize="v0", forward_type="v2", channel_first=False, k_group=1, **kwargs):
        super().__init__()
        if channel_first:
            raise NotImplementedError("channel_first=True is not used by GM-UNet")
        if act_layer is not nn.SiLU:
            raise NotImplementedError("only SiLU is fused")
        if forward_type not in ("v2", "v2_no32"):
            raise NotImplementedError(f"forward_type {forward_type!r}: only 'v2' and 'v2_no32' are on the hot path")
        if initialize != "v0":
            raise NotImplementedError("only initialize='v0'")
        d_inner = int(ssm_ratio * d_model)
        dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.d_model, self.d_inner, self.d_state, self.dt_rank, self.k_group = d_model, d_inner, d_state, dt_rank, k_group
        self.disable_force32 = forward_type.endswith("_no32")
        self.with_dconv = d_conv > 1
        self.channel_first = False
        # construction order = reference (ss2d.py:266-335): out_norm, in_proj, conv2d, x_proj, out_proj, dt_projs, A, D
        self.out_norm = nn.LayerNorm(d_inner)
        self.in_proj = nn.Linear(d_model, d_inner * 2, bias=bias)
        self.act = act_layer()
        if self.with_dconv:
            self.conv2d = nn.Conv2d(d_inner, d_inner, groups=d_inner, bias=conv_bias, kernel_size=d_conv,
                                    padding=(d_conv - 1) // 2)
        x_proj = [nn.Linear(d_inner, dt_rank + d_state * 2, bias=False) for _ in range(k_group)]
        self.x_proj_weight = nn.Parameter(torch.stack([t.weight for t in x_proj], dim=0))          # (K, R+2N, D)
        self.out_proj = nn.Linear(d_inner, d_model, bias=bias)
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else nn.Identity()
        dt_projs = [self.dt_init(dt_rank, d_inner, dt_scale, dt_init, dt_min, dt_max, dt_init_floor) for _ in range(k_group)]
        self.dt_projs_weight = nn.Parameter(torch.stack([t.weight for t in dt_projs], dim=0))      # (K, D, R)
        self.dt_projs_bias = nn.Parameter(torch.stack([t.bias for t in dt_projs], dim=0))          # (K, D)
        self.A_logs = self.A_log_init(d_state, d_inner, copies=k_group, merge=True)                # (K*D, N)
        self.Ds = self.D_init(d_inner, copies=k_group, merge=True)                                 # (K*D)

    # True: merge + out_norm + gate + out_proj as ONE tcgen05 kernel (csrc/gate_proj_tc.cu, Fn._OutGateProj) when eligible.
    # Off by default: at the north-star shape that kernel takes 170 us against 146 + 19 us for the epilogue kernel followed by
    # the tensor-core out_proj — its 96 KB operand tile allows one CTA per SM and its phases run back to back.
    fuse_out_proj = False

    # -- forward_corev2 (ss2d.py:349-500) as ONE fused pipeline ------------------------------------
    @staticmethod
    def _plan(dirs):
        """How the K directions map onto contiguous traversals. Column-major directions (2, 4) are row-major traversals
        (1, 3) of the TRANSPOSED image, so the scan kernels only ever see TMA-friendly contiguous rows; the transposition
        is paid once on x (and once on dy in the backward) and undone inside the epilogue's merge.
        -> (P input planes, per-plane transposed flags, kernel directions) or None if the pattern is not periodic."""
        tflag = [k in (2, 4) for k in dirs]
        kdirs = tuple({1: 1, 2: 1, 3: 3, 4: 3}[k] for k in dirs)
        if all(t == tflag[0] for t in tflag):
            return 1, (tflag[0],), kdirs
        if len(dirs) % 2 == 0 and all(tflag[i] == tflag[i % 2] for i in range(len(dirs))) and tflag[0] != tflag[1]:
            return 2, (tflag[0], tflag[1]), kdirs
        return None

    def forward_core(self, x: torch.Tensor, z, dirs, planes_u=None, out_proj=None):
        """x: (B, D, H, W) activated conv output; z: (B, H, W, D) raw gate view or None -> (B, H, W, D).
        planes_u: (B, 2 D, L) = [x | x transposed] already written by the convolution kernel (Fn.dwconv3_silu_planes); x is
        then only consulted for its shape. out_proj = (weight, bias): the output projection (ss2d.py:518) is applied as well —
        inside the epilogue kernel when eligible (Fn._OutGateProj) — and (B, H, W, d_model) is returned."""
        Bn, D, H, W = x.shape
        K, _, R = self.dt_projs_weight.shape
        N = self.A_logs.shape[1]
        L = H * W
        C = R + 2 * N
        assert K == len(dirs), "one direction per k_group"
        plan = self._plan(dirs)
        if plan is None:      # irregular direction sets: index-mapped traversal inside the kernels (slower, no TMA)
            P, tflags, kdirs = 1, (False,), tuple(dirs)
        else:
            P, tflags, kdirs = plan
        if planes_u is not None and not (P == 2 and tflags == (False, True)):
            planes_u = None
        Wx = self.x_proj_weight.to(planes_u.dtype if planes_u is not None else x.dtype)
        planes = None if planes_u is not None else [x.transpose(2, 3).reshape(Bn, D, L) if t else x.reshape(Bn, D, L) for t in tflags]
        # pointwise projections evaluated in each plane's own pixel order (identical values to projecting the permuted xs)
        if planes_u is not None:
            # both planes come from the convolution kernel in ONE buffer: one GEMM with block-structured weights (direction k
            # reads plane k % 2) gives x_dbl in group order — no transposed copy, no cat, no stack
            u = planes_u
            W_blk = torch.cat([F.pad(Wx[k], ((k % 2) * D, (1 - k % 2) * D)) for k in range(K)])         # (K C, 2 D)
            x_dbl = Fn.proj_cm(W_blk, u).view(Bn, K, C, L)
        elif P == 1:
            u = planes[0]
            x_dbl = Fn.proj_cm(Wx.reshape(K * C, D), u).view(Bn, K, C, L)
        else:
            u = torch.cat(planes, dim=1)                                                     # (B, 2D, L)
            per_plane = [Fn.proj_cm(Wx[j::2].reshape((K // 2) * C, D), planes[j]).view(Bn, K // 2, C, L) for j in range(2)]
            x_dbl = torch.stack(per_plane, dim=2).view(Bn, K, C, L)                          # group order restored
        dts_r, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
        # dt_proj (ss2d.py:469). The scan takes delta in fp32 (force_fp32, :479-480): under autocast the R = dt_rank rows are
        # cast up (a few MB) and the projection runs in fp32, so delta — the largest tensor of the call — is written once, in
        # the type the scan reads, instead of once in bf16, read again and written again in fp32.
        up = (not self.disable_force32) and dts_r.dtype != torch.float32
        with torch.autocast("cuda", enabled=not up and torch.is_autocast_enabled("cuda")):
            w_dt = self.dt_projs_weight.float() if up else self.dt_projs_weight.to(dts_r.dtype)
            r_in = dts_r.float() if up else dts_r
            if K == 1:
                dts = Fn.proj_cm(w_dt[0], r_in[:, 0])                                        # (B, D, L)
            else:
                dts = torch.matmul(w_dt.unsqueeze(0), r_in).reshape(Bn, K * D, L)
        As = -torch.exp(self.A_logs.float())
        Ds = self.Ds.float()
        bias = self.dt_projs_bias.reshape(-1).float()
        if not self.disable_force32:
            u, dts, Bs, Cs = u.float(), dts.float(), Bs.float(), Cs.float()
        else:
            dts, Bs, Cs = dts.to(u.dtype), Bs.to(u.dtype), Cs.to(u.dtype)
        ys = Fn._SS2DScanNatural.apply(u.contiguous(), dts, As, Bs, Cs, Ds, bias, H, W, kdirs, P)   # (B, K/P, P, D, L)
        zz = z.reshape(Bn, L, D) if z is not None else None
        if D > Fn.ops.out_gate_max_D(backward=torch.is_grad_enabled()):
            # rows wider than the fused epilogue's shared-memory tiles (VMamba stage 4: d_inner = 1536): merge, LayerNorm and
            # gate as separate passes, exactly the reference's sequence (ss2d.py:486-498, 515-517)
            planes_y = [ys[:, i, j].reshape(Bn, D, W, H).transpose(2, 3).reshape(Bn, D, L) if tflags[j] else ys[:, i, j]
                        for i in range(K // P) for j in range(P)]
            if len(planes_y) == 4:
                ym = (planes_y[0] + planes_y[2]) + (planes_y[1] + planes_y[3])                   # csms6s.py:38-39
            else:
                ym = planes_y[0]
                for q in planes_y[1:]:
                    ym = ym + q
            y = F.layer_norm(ym.transpose(1, 2), (D,), self.out_norm.weight.float(), self.out_norm.bias.float(),
                             self.out_norm.eps).to(x.dtype)
            if zz is not None:
                y = y * F.silu(zz)
            y = y.view(Bn, H, W, D)
            return y if out_proj is None else Fn.linear_tc(y, out_proj[0], out_proj[1])
        tplanes = sum(1 << j for j, t in enumerate(tflags) if t)
        if out_proj is not None and self.fuse_out_proj and Fn.out_gate_proj_ok(ys, zz, out_proj[0], H, W):
            # merge + out_norm + gate + out_proj in one tensor-core kernel: the gated tensor stays in shared memory
            out = Fn._OutGateProj.apply(ys, self.out_norm.weight.float(), self.out_norm.bias.float(), zz, out_proj[0], out_proj[1],
                                        self.out_norm.eps, H, W, tplanes)
            return out.view(Bn, H, W, -1)
        y = Fn._OutGate.apply(ys, self.out_norm.weight.float(), self.out_norm.bias.float(), zz, True,
                              self.out_norm.eps, x.dtype, H, W, tplanes).view(Bn, H, W, D)
        return y if out_proj is None else Fn.linear_tc(y, out_proj[0], out_proj[1])

    def forward(self, x: torch.Tensor, CrossScan=None, CrossMerge=None, **kwargs):
        if CrossScan is None and CrossMerge is None:
            dirs = (1, 2, 3, 4) if self.k_group == 4 else tuple(range(1, self.k_group + 1))
        else:
            dirs = Fn.directions_of(CrossScan, CrossMerge)
            if dirs is None:
                raise NotImplementedError("SS2D.forward needs a matching (CrossScan*, CrossMerge*) pair")
        pz = Fn.in_proj_planes(x, self.in_proj.weight, self.in_proj.bias)    # ss2d.py:504-510 in one tensor-core launch
        if pz is not None:
            xi, z = pz
        else:
            xz = Fn.linear_ts(x, self.in_proj.weight, self.in_proj.bias)    # ss2d.py:504 (tall-skinny weight gradient)
            xi, z = xz.chunk(2, dim=-1)                               # :506  (views; SiLU(z) is fused into the epilogue)
            xi = xi.permute(0, 3, 1, 2).contiguous()                  # :510
        cv = self.conv2d if self.with_dconv else None
        planes_u = None
        if (cv is not None and isinstance(self.act, nn.SiLU) and cv.kernel_size == (3, 3) and cv.padding == (1, 1)
                and cv.stride == (1, 1) and cv.dilation == (1, 1) and cv.groups == cv.in_channels == cv.out_channels
                and cv.padding_mode == "zeros" and xi.is_cuda):
            plan = self._plan(dirs)
            if plan is not None and plan[0] == 2 and plan[1] == (False, True):
                # K = 4 (directions 1, 2, 3, 4): the convolution kernel writes the natural AND the transposed plane
                planes_u = Fn.dwconv3_silu_planes(xi, cv.weight, cv.bias)      # :512-513 + CrossScan's layout, one kernel
            if planes_u is None:
                xi = Fn.dwconv3_silu(xi, cv.weight, cv.bias)      # :512-513 as one kernel
        else:
            if cv is not None:
                xi = Fn.dwconv3(xi, cv)                           # :512 (reduction-shaped parameter gradient)
            xi = self.act(xi)                                     # :513
        y = self.forward_core(xi, z, dirs, planes_u, (self.out_proj.weight, self.out_proj.bias))    # :514-518
        return self.dropout(y)


class GroupMambaLayer(nn.Module):
    """Four single-direction SS2Ds on channel quarters + channel-affinity gating (groupmamba.py:85-159)."""

    def __init__(self, input_dim, output_dim, d_state=1, d_conv=3, expand=1, reduction=16):
        super().__init__()
        reduced = input_dim // reduction
        self.fc1 = nn.Linear(input_dim, reduced, bias=True)
        self.fc2 = nn.Linear(reduced, output_dim, bias=True)
        self.relu = nn.ReLU()
        self.sigmoid = nn.Sigmoid()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.norm = nn.LayerNorm(input_dim)
        for i in range(1, 5):
            setattr(self, f"mamba_g{i}", SS2D(d_model=input_dim // 4, d_state=d_state, ssm_ratio=expand, d_conv=d_conv))
        self.proj = nn.Linear(input_dim, output_dim)
        self.skip_scale = nn.Parameter(torch.ones(1))

    # plane p of the grouped tensors holds group _PLANES[p]: the two row-major directions first (natural pixel order), then the
    # two column-major ones (stored transposed, where they are row-major too); the permutation is its own inverse
    _PLANES = (0, 2, 1, 3)
    grouped = True          # False: run the four SS2D modules one after the other (the reference's own sequence of calls)

    def _can_group(self, x, H, W) -> bool:
        ms = [self.mamba_g1, self.mamba_g2, self.mamba_g3, self.mamba_g4]
        m0 = ms[0]
        return (self.grouped and H == W and x.is_cuda and m0.with_dconv and m0.k_group == 1 and not m0.disable_force32
                and m0.in_proj.bias is None and m0.out_proj.bias is None and isinstance(m0.dropout, nn.Identity)
                and m0.conv2d.kernel_size == (3, 3) and m0.conv2d.bias is not None
                and m0.d_inner <= Fn.ops.out_gate_max_D(backward=torch.is_grad_enabled()))

    def _ss2d_grouped(self, xn, H, W):
        """The four single-direction SS2Ds (ss2d.py:502-519, called at groupmamba.py:143-146) as ONE pipeline: block-structured
        projections, one depthwise convolution, one scan launch (G = 4, one direction per group), one epilogue launch that
        also concatenates. xn: (B, L, C) normalised input -> (B, L, C) = cat_g SS2D_g(xn_g)."""
        ms = [self.mamba_g1, self.mamba_g2, self.mamba_g3, self.mamba_g4]
        PL = self._PLANES
        Bn, L, C = xn.shape
        D, N, R = ms[0].d_inner, ms[0].d_state, ms[0].dt_rank
        # in_proj of the four groups (ss2d.py:504): x part with its output blocks in plane order, z part in group order
        w_in = [m.in_proj.weight for m in ms]                                       # (2 D, C / 4) each
        Cq = C // 4
        Wx = torch.cat([F.pad(w_in[g][:D], (g * Cq, (3 - g) * Cq)) for g in PL])    # row block p = group PL[p], its own input columns
        Wz = torch.block_diag(*[w[D:] for w in w_in])
        x_all, z_all = Fn._InProjSplit.apply(xn, Wx, Wz)                            # (B, L, 4 D) each
        u = Fn._ToPlanes.apply(x_all.view(Bn, H, W, 4 * D), 2 * D)                  # (B, 4 D, H, W), planes 2, 3 transposed (:510)
        # depthwise 3 x 3 (:512): a transposed image is convolved with the transposed kernel
        Wc = torch.cat([ms[PL[0]].conv2d.weight, ms[PL[1]].conv2d.weight,
                        ms[PL[2]].conv2d.weight.transpose(2, 3), ms[PL[3]].conv2d.weight.transpose(2, 3)])
        bc = torch.cat([ms[g].conv2d.bias for g in PL])
        u = Fn.dwconv3_silu(u, Wc, bc).view(Bn, 4 * D, L)                            # :512-513, one kernel
        # x_proj / dt_proj (:465-469) as two block-structured GEMMs from u: delta = (W_dt W_x[:R]) u and [B; C] = W_x[R:] u
        wxp = [ms[g].x_proj_weight[0] for g in PL]                                  # (R + 2 N, D)
        W_dt = torch.block_diag(*[torch.matmul(ms[g].dt_projs_weight[0], wx[:R]) for g, wx in zip(PL, wxp)])     # (4 D, 4 D)
        W_bc = torch.block_diag(*[wx[R:] for wx in wxp])                            # (4 * 2 N, 4 D)
        if u.dtype != torch.float32:
            # delta is consumed in fp32 by the scan (:479-480): project the fp32 copy of u the scan reads anyway, so delta is
            # written once in that type (no bf16 round trip + cast of the largest tensor of the call)
            uf = u.float()
            with torch.autocast("cuda", enabled=False):
                dts = Fn.proj_cm(W_dt.float(), uf)                                  # (B, 4 D, L) fp32
        else:
            uf = u
            dts = Fn.proj_cm(W_dt.to(u.dtype), u)                                   # (B, 4 D, L)
        BC = Fn.proj_cm(W_bc.to(u.dtype), u).view(Bn, 4, 2 * N, L)
        As = -torch.exp(torch.cat([ms[g].A_logs for g in PL]).float())              # :473
        Ds = torch.cat([ms[g].Ds for g in PL]).float()
        bias = torch.cat([ms[g].dt_projs_bias.reshape(-1) for g in PL]).float()
        uf, dts, BC = uf.contiguous(), dts.float(), BC.float()                      # force_fp32 (:479-480)
        ys = Fn._SS2DScanNatural.apply(uf, dts, As, BC[:, :, :N], BC[:, :, N:], Ds, bias, H, W, (1, 3, 1, 3), 4)
        lnw = torch.stack([m.out_norm.weight for m in ms]).float()
        lnb = torch.stack([m.out_norm.bias for m in ms]).float()
        y = Fn._GroupGate.apply(ys.view(Bn, 4, D, L), lnw, lnb, z_all.contiguous(), ms[0].out_norm.eps, u.dtype, H, W, PL, 0b1100)
        W_out = torch.block_diag(*[m.out_proj.weight for m in ms])                  # (C, 4 D)  (:518)
        return Fn.linear_ts(y, W_out, None)

    def forward(self, x, H, W):
        if x.dtype == torch.float16:
            x = x.type(torch.float32)
        Bn, L, C = x.shape
        x = Fn.layer_norm_rows(x, self.norm.weight, self.norm.bias, self.norm.eps)   # :131 (row-wise LN kernel)
        aff = self.sigmoid(self.fc2(self.relu(self.fc1(x.mean(dim=1)))))        # :134-137 channel affinity
        x4 = x.view(Bn, H, W, C)
        if self._can_group(x, H, W):
            ycat = self._ss2d_grouped(x, H, W).view(Bn, H, W, C)
        else:
            parts = torch.chunk(x4, 4, dim=-1)
            pairs = ((Fn.CrossScan_1, Fn.CrossMerge_1), (Fn.CrossScan_2, Fn.CrossMerge_2),
                     (Fn.CrossScan_3, Fn.CrossMerge_3), (Fn.CrossScan_4, Fn.CrossMerge_4))
            outs = [getattr(self, f"mamba_g{i + 1}")(parts[i], CrossScan=pairs[i][0], CrossMerge=pairs[i][1]) for i in range(4)]
            ycat = torch.cat(outs, dim=-1)
        xm = ycat * self.skip_scale * x4                                        # :149
        xm = xm.view(Bn, L, C) * aff.unsqueeze(1)                               # :154
        xm = Fn.layer_norm_rows(xm, self.norm.weight, self.norm.bias, self.norm.eps)   # :156 (same LayerNorm, shared weights)
        return Fn.linear_ts(xm, self.proj.weight, self.proj.bias)               # :157


class LayerNormRows(nn.LayerNorm):
    """nn.LayerNorm over the last dimension on this package's row kernel (csrc/layernorm.cu, one warp per row): the norm in
    front of the FFN in the reference's Block_mamba (`self.norm2`, groupmamba.py:200, 225) and the patch-embedding norms.
    Same parameters and state_dict as nn.LayerNorm (an existing instance can be re-classed in place); rows wider than the
    kernel's 512 channels, inputs without affine parameters and CPU tensors keep nn.LayerNorm's own forward."""

    def forward(self, x):
        if (x.is_cuda and x.dim() >= 2 and len(self.normalized_shape) == 1 and self.elementwise_affine
                and self.normalized_shape[0] <= Fn.ops.LN_MAX_C):
            return Fn.layer_norm_rows(x, self.weight, self.bias, self.eps)
        return super().forward(x)


# ---- the GroupMamba FFNs: same constructors, sub-module names and initialisation as the reference, the
#      depthwise stack between the two linear layers on this package's channels-last kernels ----
def _ffn_init_weights(m):
    """PVT2FFN._init_weights / custom_ffn._init_weights (groupmamba.py:63-76, custom_mlp.py:349-361)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
        fan_out //= m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


class DWConv(nn.Module):
    """Parameter holder of the reference's DWConv (groupmamba.py:446-455): `dwconv` = depthwise 3 x 3 with bias."""

    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)


class InceptionDWConv2d_MultiScale(nn.Module):
    """Parameter holder of custom_mlp.py:313-336: identity | 3 x 3 | 5 x 5 | 7 x 7 depthwise branches on channel segments."""

    def __init__(self, in_channels, kernel_sizes=(1, 3, 5), branch_ratio=0.125):
        super().__init__()
        gc = int(in_channels * branch_ratio)
        self.dwconv_3x3 = nn.Conv2d(gc, gc, kernel_size=3, padding=1, groups=gc)
        self.dwconv_5x5 = nn.Conv2d(gc, gc, kernel_size=5, padding=2, groups=gc)
        self.dwconv_7x7 = nn.Conv2d(gc, gc, kernel_size=7, padding=3, groups=gc)
        self.split_indexes = (in_channels - 3 * gc, gc, gc, gc)


class PVT2FFN(nn.Module):
    """fc1 -> depthwise 3 x 3 -> GELU -> fc2 (groupmamba.py:54-83); forward(x (B, L, C), H, W)."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.apply(_ffn_init_weights)

    def forward(self, x, H, W):
        h = self.fc1(x)
        y = Fn.ffn_depthwise(h, (H, W), self.dwconv.dwconv)
        return self.fc2(y)


class custom_ffn(nn.Module):
    """fc1 -> depthwise 3 x 3 -> GELU -> multi-scale depthwise residual -> fc2 (custom_mlp.py:338-368)."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.act = nn.GELU()
        self.custom = InceptionDWConv2d_MultiScale(hidden_features, [])
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.apply(_ffn_init_weights)

    def forward(self, x, H, W):
        h = self.fc1(x)
        c = self.custom
        y = Fn.ffn_depthwise(h, (H, W), self.dwconv.dwconv, (c.dwconv_3x3, c.dwconv_5x5, c.dwconv_7x7))
        return self.fc2(y)
