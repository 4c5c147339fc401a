"""MixLinear_GEMM — the mixed-precision quantized Linear (W8A8O16 / W4A4O16) on the B200 fused kernel.

Host-side mirror of /root/reference/mixquant/modules/linear.py:26-376 (same class name, constructor,
`from_linear`, `forward(x, cache=None, unfused=False)`, `forward_without_preconditionFusedSilu(x, cache)`,
`FindOutliers`, same state: q_weight / scale_col / ind / weight_cache / bias / cnt / add_outliers), written
against libmixq_sm100's C ABI instead of the `mixlib` + `torch.mm` call chain:

  reference steady state (linear.py:187-193, :244-283)      here
  ---------------------------------------------------      -----------------------------------------
  ExtractOutliersAndSetToZeros + FindRowScale               \
  torch.mm(activation_outliers, weight_cache.T)              >  ONE launch: mixq_linear_fused
  int8FusedDequantize[Silu] (+ y1 += bias)                  /

The online outlier discovery of the first `cache.stop` calls (linear.py:200-226) keeps the reference's
sequence — and its one host synchronisation per call — but uses the library's scan/compaction kernels
instead of torch.where/torch.unique.

There is no CPU path and no PyTorch fallback: `forward` raises if the tensors are not CUDA tensors or the
shared library is missing.  The weight quantisation in `from_linear` is offline torch arithmetic (exactly
the reference's expressions) and runs on whatever device the weights are on.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .cache import MixLibCache, _PAD, _round_up

ACT_NONE, ACT_SILU = 0, 1


def two_compl(x: torch.Tensor, bits: int) -> torch.Tensor:
    return torch.where(x < 0, 2 ** bits + x, x)


def pack_to_i4(X: torch.Tensor) -> torch.Tensor:
    """linear.py:12-18: two's-complement nibbles, low nibble = even column, high nibble = odd column."""
    X_i8 = two_compl(X.to(dtype=torch.int8), 4).to(torch.uint8)
    return X_i8[:, 0::2] | (X_i8[:, 1::2] << 4)


def unpack_int8_to_int4(weight, ind):
    """linear.py:20-22 (the reference's name, kept): fp16 [N, len(ind)] un-scaled nibble values."""
    from . import mixlib
    assert weight.dim() == 2
    return mixlib.unpack_int4_to_fp16(weight, ind)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class MixLinear_GEMM(nn.Module):
    def __init__(self, in_features, out_features, bias, dev, bit, weight_only=False, cache=None,
                 fp_features_num=128, name=None):
        super().__init__()
        if weight_only:
            # EETQ W8A16 weight-only path (linear.py:58-61, :178-184): never taken for Llama
            # (utils/module.py:6 weight_only_map["LlamaForCausalLM"] = []) — outside the hot path.
            raise NotImplementedError("weight_only (EETQ W8A16) is outside the MixLinear hot path")
        if bit not in (4, 8):
            raise ValueError("bit must be 8 or 4")
        self.in_features = in_features
        self.out_features = out_features
        self.bit = bit
        self.register_buffer("scale_col", torch.empty((1, out_features), dtype=torch.float16, device=dev))
        if bit == 8:
            self.register_buffer("q_weight", torch.empty((out_features, in_features), dtype=torch.int8, device=dev))
            n0 = 0
        else:
            self.fp_features_num = fp_features_num
            self.register_buffer("q_weight", torch.empty((out_features, in_features // 2), dtype=torch.uint8, device=dev))
            n0 = fp_features_num
        # outlier state: capacity buffers + a count; `ind` / `weight_cache` are the reference-shaped views
        self._n_ind = n0
        self._ind_buf = torch.zeros(in_features, dtype=torch.int32, device=dev)
        cap = max(_PAD, _round_up(n0, _PAD))
        self._wc_buf = torch.zeros((out_features, cap), dtype=torch.float16, device=dev)
        if bias:
            self.register_buffer("bias", torch.empty((out_features,), dtype=torch.float16, device=dev))
        else:
            self.bias = None
        self.cnt = 0
        self.forward_without_precondition_len = -1 if bit == 8 else fp_features_num
        self.cache = cache
        self.weight_only = False
        self.add_outliers = True
        if cache is not None:
            self.sigma = torch.ones((1, 1), dtype=torch.float16, device=dev)
            self.sigma[0] = cache.sigma.reshape(-1)[0]
        self._sigma_f = float(cache.sigma.reshape(-1)[0]) if cache is not None else 6.0   # host copy: no sync per launch
        self.arch = 10  # sm_100a only; the reference's arch == 9 split path (linear.py:234-241) is never taken
        self.name = name
        self._args = _lib.LinearArgs()

    # ------------------------------------------------------------------ reference-shaped state
    @property
    def ind(self) -> torch.Tensor:
        return self._ind_buf[: self._n_ind]

    @ind.setter
    def ind(self, value: torch.Tensor):
        n = int(value.shape[0])
        if n > self._ind_buf.shape[0]:
            raise ValueError("more outlier columns than input features")
        self._ind_buf[:n] = value.to(self._ind_buf.device, torch.int32)
        self._n_ind = n

    @property
    def weight_cache(self):
        if self._n_ind == 0:
            return None
        return self._wc_buf[:, : self._n_ind]

    @weight_cache.setter
    def weight_cache(self, value):
        if value is None:
            return
        n = int(value.shape[1])
        self._reserve_wc(n)
        self._wc_buf[:, :n] = value.to(self._wc_buf.device, torch.float16)

    def _reserve_wc(self, n: int):
        if n > self._wc_buf.shape[1]:
            new = torch.zeros((self.out_features, _round_up(n, _PAD)), dtype=torch.float16, device=self._wc_buf.device)
            new[:, : self._wc_buf.shape[1]] = self._wc_buf
            self._wc_buf = new

    # bit 4: the reference registers `weight_cache` [N, fp_features_num] and `ind` [fp_features_num] as buffers
    # (linear.py:49-57), so they are state_dict keys that base.py's load_checkpoint_in_model sets by name.  Here they are
    # views of the capacity buffers; the two hooks below keep the reference's keys, dtypes and shapes in state_dict().
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        if self.bit == 4:
            n = self.fp_features_num
            destination[prefix + "weight_cache"] = self._wc_buf[:, :n].detach().clone()
            destination[prefix + "ind"] = self._ind_buf[:n].detach().clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        if self.bit == 4:
            n = self.fp_features_num
            wc, ind = state_dict.get(prefix + "weight_cache"), state_dict.get(prefix + "ind")
            for key, val, shape in ((prefix + "weight_cache", wc, (self.out_features, n)), (prefix + "ind", ind, (n,))):
                if val is None:
                    if strict:
                        missing_keys.append(key)
                elif tuple(val.shape) != shape:
                    error_msgs.append(f"size mismatch for {key}: checkpoint {tuple(val.shape)}, module {shape}")
            if wc is not None and ind is not None and tuple(wc.shape) == (self.out_features, n) and tuple(ind.shape) == (n,):
                self._reserve_wc(n)
                self._wc_buf[:, :n] = wc.to(self._wc_buf.device, torch.float16)
                self._ind_buf[:n] = ind.to(self._ind_buf.device, torch.int32)
                self._n_ind = n
            # the base class must not report the two keys as unexpected
            state_dict = {k: v for k, v in state_dict.items() if k not in (prefix + "weight_cache", prefix + "ind")}
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._ind_buf = fn(self._ind_buf)
        self._wc_buf = fn(self._wc_buf)
        if hasattr(self, "sigma"):
            self.sigma = fn(self.sigma)
        return self

    # ------------------------------------------------------------------ offline weight quantisation
    @classmethod
    def from_linear(cls, linear, bit, weight_only=False, init_only=False, cache=None, layer_scales=None,
                    dev="cuda", name=None, fp_features_num=128):
        """linear.py:88-150.  `linear` is an nn.Linear (or anything with .weight/.bias/.in_features/.out_features)."""
        q = cls(linear.in_features, linear.out_features, linear.bias is not None, dev, bit=bit,
                weight_only=weight_only, cache=cache, name=name, fp_features_num=fp_features_num)
        if init_only:
            return q
        w = linear.weight.data
        if bit == 8:
            # linear.py:111-119
            scale = (torch.max(torch.abs(w), dim=1)[0].unsqueeze(1) / 127).to(torch.float16).reshape((1, linear.out_features))
            q.scale_col.copy_(scale)
            tmp = w.to(dev).clone()
            tmp /= q.scale_col.T.to(tmp.dtype)
            q.q_weight.copy_(tmp.round().to(torch.int8))
        else:
            # linear.py:121-143
            assert layer_scales is not None
            ind = torch.sort(layer_scales)[1][-fp_features_num:]
            tmp = w.to(dev).clone()
            q._wc_buf[:, :fp_features_num] = tmp[:, ind.to(tmp.device)].to(torch.float16)
            tmp[:, ind.to(tmp.device)] = 0
            scale = (torch.max(torch.abs(tmp), dim=1)[0].unsqueeze(1) / 10).to(torch.float16).reshape((1, linear.out_features))
            q.scale_col.copy_(scale)
            tmp /= q.scale_col.T.to(tmp.dtype)
            tmp = torch.clamp(tmp.round(), -8, 7)
            q.q_weight.copy_(pack_to_i4(tmp.to(torch.int8)).to(dev))
            q._ind_buf[:fp_features_num] = ind.to(dev, torch.int32)
            q._n_ind = fp_features_num
        if linear.bias is not None:
            q.bias.copy_(linear.bias.half())
        return q

    # ------------------------------------------------------------------ helpers
    @torch.no_grad()
    def FindOutliers(self, Activation):
        """linear.py:157-161 — sorted unique ids of the columns holding any |x| > sigma (int32).
        Device-side: the fused scan flags the columns, one compaction kernel orders them."""
        cache = self.cache
        x2 = Activation.reshape(-1, Activation.shape[-1])
        M, K = x2.shape
        self._scan(cache, x2, M)
        tmp = torch.empty(K, dtype=torch.int32, device=x2.device)
        n_new = self._compact(cache, 0, out=tmp)
        return tmp[:n_new]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _require_cuda(self, *ts):
        for t in ts:
            if t is not None and not t.is_cuda:
                raise _lib.MixqError("MixLinear_GEMM.forward needs CUDA tensors: there is no CPU path")

    def _scan(self, cache, inputs, M):
        """FindRowScale + threshold test + outlier-column flags in one launch (linear.py:190-193, :201, :203)."""
        K = self.in_features
        q_x = cache.q_x_buffer(M, K)
        col_over = cache.col_over_buffer(K)
        cache.over_flag.zero_()
        # a row with amax one fp16 ulp above sigma sets its column flags without raising over_flag (fp16(amax/qmax) ==
        # fp16(sigma/qmax)): flags nobody compacted must not leak into the next module's discovery through the shared cache
        col_over.zero_()
        lib = _lib.load()
        _lib.check(lib.mixq_find_row_scale_scan(_ptr(inputs), _ptr(cache.x_scale), _ptr(q_x), M, K, self.bit,
                                                self._sigma_f, _ptr(col_over), _ptr(cache.over_flag),
                                                self._stream()), "FindRowScale(scan)")
        cache.q_xcache = q_x

    def _compact(self, cache, n_have, out=None):
        K = self.in_features
        dst = self._ind_buf if out is None else out
        lib = _lib.load()
        _lib.check(lib.mixq_compact_outlier_columns(_ptr(cache.col_over_buffer(K)), K, dst.data_ptr() + 4 * n_have,
                                                    dst.shape[0] - n_have, _ptr(cache.n_new), self._stream()),
                   "compact_outlier_columns")
        return int(cache.n_new.item())  # host sync — discovery calls only, like linear.py:201

    def _gather_weight_columns(self, ind_new: torch.Tensor, col0: int):
        """weight_cache[:, col0:col0+n] = q_weight[:, ind].half() * scale_col.T (linear.py:207, :209-210)."""
        n = int(ind_new.shape[0])
        self._reserve_wc(col0 + n)
        lib = _lib.load()
        _lib.check(lib.mixq_gather_weight_columns(_ptr(self.q_weight), _ptr(self.scale_col), _ptr(ind_new), n,
                                                  _ptr(self._wc_buf), self._wc_buf.shape[1], col0, self.out_features,
                                                  self.in_features, self.bit, self._stream()), "gather_weight_columns")

    def _launch(self, cache, M, y, *, x=None, skip_prologue=False, act=ACT_NONE, norm_weight=None, eps=0.0,
                norm_out=None, residual=None, q_x=None, act_outliers=None, ld_ao=None, up=None, push=None):
        """One mixq_linear_fused launch."""
        a = self._args
        n = self._n_ind
        if q_x is None:
            q_x = cache.q_x_buffer(M, self.in_features)
        if act_outliers is None:
            ao = cache.ao_buffer(n)
            act_outliers, ld_ao = ao, ao.shape[1]
        a.x = _ptr(x)
        a.norm_weight = _ptr(norm_weight)
        a.norm_out = _ptr(norm_out)
        a.eps = float(eps)
        a.M, a.N, a.K = M, self.out_features, self.in_features
        a.q_weight = _ptr(self.q_weight)
        a.scale_col = _ptr(self.scale_col)
        a.bias = _ptr(self.bias)
        a.bit = self.bit
        # SwiGLU pair: self is gate_proj, `up` is up_proj (same outlier set)
        a.q_weight_up = _ptr(None if up is None else up.q_weight)
        a.scale_col_up = _ptr(None if up is None else up.scale_col)
        a.weight_cache_up = _ptr(None if up is None else up._wc_buf)
        a.ind = _ptr(self._ind_buf)
        a.n_ind = n
        a.weight_cache = _ptr(self._wc_buf)
        a.ld_wc = self._wc_buf.shape[1]
        a.q_x = _ptr(q_x)
        a.x_scale = _ptr(cache.x_scale)
        a.act_outliers = _ptr(act_outliers)
        a.ld_ao = ld_ao
        a.sigma = self._sigma_f
        a.col_over = 0
        a.over_flag = 0
        a.residual = _ptr(residual)
        a.ld_res = 0 if residual is None else residual.stride(0)
        a.y = _ptr(y)
        # tensor-parallel push: column slice j of y goes straight into rank j's receive slot (tp.PushExchange.push_targets())
        if push is None:
            a.peer_cols = a.peer_bcast = 0
        else:
            ptrs, a.peer_cols, a.peer_bcast = push
            for j, pj in enumerate(ptrs):
                a.y_peer[j] = pj
        if M <= 128 and push is None:      # few tiles: let the library split K over several CTAs per tile
            ws = cache.splitk_buffer()
            a.splitk_ws, a.splitk_ws_bytes = ws.data_ptr(), ws.numel() * 4
        else:
            a.splitk_ws, a.splitk_ws_bytes = 0, 0
        a.act = act
        a.skip_prologue = 1 if skip_prologue else 0
        a.grid_sync = _ptr(cache.grid_sync)
        a.tile_n = 0
        _lib.check(_lib.load().mixq_linear_fused(C.byref(a), self._stream()), f"mixq_linear_fused({self.name})")

    def _cached_act_outliers(self, cache, M):
        """Fused mode: the producer (RMSNorm) left activation_outliers somewhere; the kernel's TMA needs a
        16-byte aligned base and row pitch."""
        n = self._n_ind
        ao = cache.activation_outliers
        if n == 0 or ao is None:
            return None, None
        if ao.stride(1) == 1 and ao.stride(0) % 8 == 0 and ao.data_ptr() % 16 == 0:
            return ao, ao.stride(0)
        buf = cache.ao_buffer(n)
        buf[:M, :n] = ao
        return buf, buf.shape[1]

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, cache=None, unfused=False, residual=None, out=None, push=None):
        """linear.py:165-289.  unfused=True: x holds raw fp16 activations (their outlier columns are zeroed IN
        PLACE, as the reference does); unfused=False: the preceding FasterTransformerRMSNorm already left
        q_xcache / x_scale / activation_outliers in the cache.  `residual` (extension): y = fp16(y + residual);
        `out` (extension): caller-owned contiguous fp16 [M, N] to write y into (e.g. a peer-mapped exchange buffer)."""
        if cache is None:
            cache = self.cache
        cache.shape = x.shape[:-1] + (self.out_features,)
        inputs = x.reshape(-1, x.shape[-1])
        self._require_cuda(inputs, self.q_weight)
        if unfused and inputs.data_ptr() != x.data_ptr():
            raise _lib.MixqError("unfused MixLinear needs a contiguous activation tensor (columns are zeroed in place)")
        M = inputs.shape[0]
        if push is not None:
            # `push` (extension): (slot pointers, slice width) from tp.PushExchange — the fp16 output is scattered straight
            # into the ranks' receive slots by the epilogue; nothing is returned
            if residual is not None or out is not None or self.bias is not None:
                raise _lib.MixqError("push excludes residual / out / bias (the exchange's finish kernel adds the residual)")
            y = None
        elif out is None:
            y = torch.empty((M, self.out_features), dtype=torch.float16, device=inputs.device)
        else:
            if out.dtype != torch.float16 or not out.is_contiguous() or out.numel() != M * self.out_features:
                raise _lib.MixqError("out must be a contiguous fp16 tensor of M * out_features elements")
            y = out.view(M, self.out_features)
        res2 = None if residual is None else residual.reshape(M, self.out_features)

        if not self.add_outliers:
            cache.ind = self.ind
            if unfused:
                # steady state: gather+zero, absmax, quantise, both GEMMs, epilogue — one launch
                self._launch(cache, M, y, x=inputs, residual=res2, push=push)
                cache.q_xcache = cache.q_x_buffer(M, self.in_features)
                cache.activation_outliers = cache.ao_buffer(self._n_ind)[:M, : self._n_ind]
            else:
                ao, ld = self._cached_act_outliers(cache, M)
                self._launch(cache, M, y, skip_prologue=True, residual=res2, q_x=cache.q_xcache, act_outliers=ao, ld_ao=ld, push=push)
            return None if y is None else y.reshape(cache.shape)

        # ---- online outlier discovery (first cache.stop calls): linear.py:187-226
        lib = _lib.load()
        n = self._n_ind
        ao = cache.ao_buffer(n)
        if unfused:
            if n:
                _lib.check(lib.mixq_extract_outliers_and_set_to_zeros(_ptr(self._ind_buf), n, _ptr(inputs), _ptr(ao),
                                                                      ao.shape[1], M, self.in_features, self._stream()),
                           "ExtractOutliersAndSetToZeros")
        elif n:
            src, _ = self._cached_act_outliers(cache, M)
            if src is not None and src.data_ptr() != ao.data_ptr():
                ao[:M, :n] = src[:M, :n]
        self._scan(cache, inputs, M)   # == FindRowScale; in fused mode it recomputes the same q_x / x_scale
        cache.ind = self.ind
        if bool(cache.over_flag.item()):   # x_scale[:M].max() > sigma / qmax  (linear.py:201), host sync
            n_new = self._compact(cache, n)
            if n_new:
                ind_new = self._ind_buf[n : n + n_new]
                cache.new_ind = ind_new.clone()
                ao = cache.ao_buffer(n + n_new)
                _lib.check(lib.mixq_extract_outliers_and_set_to_zeros(_ptr(ind_new), n_new, _ptr(inputs),
                                                                      ao.data_ptr() + 2 * n, ao.shape[1], M,
                                                                      self.in_features, self._stream()),
                           "ExtractOutliersAndSetToZeros(new)")
                self._gather_weight_columns(ind_new, n)
                self._n_ind = n + n_new
                cache.ind = self.ind
                _lib.check(lib.mixq_find_row_scale(_ptr(inputs), _ptr(cache.x_scale), _ptr(cache.q_xcache), M,
                                                   self.in_features, self.bit, self._stream()), "FindRowScale")
        self.cnt += 1
        if self.cnt >= self.cache.stop or self._n_ind > 128:
            self.add_outliers = False
        cache.activation_outliers = ao[:M, : self._n_ind]
        self._launch(cache, M, y, skip_prologue=True, residual=res2, push=push)
        return None if y is None else y.reshape(cache.shape)

    @torch.no_grad()
    def forward_quantized(self, M, cache=None, residual=None, out=None, push=None):
        """The reference's fused call mode (linear.py:194-199: consume cache.q_xcache / x_scale / activation_outliers left by
        a producer kernel) when the producer keeps no fp16 activation tensor at all — e.g. the attention kernel that
        quantises its own output rows for o_proj (mixq_rope_attention_decode_quant).  Steady state only."""
        if cache is None:
            cache = self.cache
        if self.add_outliers:
            raise _lib.MixqError("forward_quantized is a steady-state path: run the discovery calls first")
        self._require_cuda(cache.q_xcache, self.q_weight)
        if push is not None:
            y = None
        elif out is None:
            y = torch.empty((M, self.out_features), dtype=torch.float16, device=self.q_weight.device)
        else:
            if out.dtype != torch.float16 or not out.is_contiguous() or out.numel() != M * self.out_features:
                raise _lib.MixqError("out must be a contiguous fp16 tensor of M * out_features elements")
            y = out.view(M, self.out_features)
        cache.shape = (M, self.out_features)
        cache.ind = self.ind
        ao, ld = self._cached_act_outliers(cache, M)
        self._launch(cache, M, y, skip_prologue=True, residual=None if residual is None else residual.reshape(M, self.out_features),
                     q_x=cache.q_xcache, act_outliers=ao, ld_ao=ld, push=push)
        return y

    @torch.no_grad()
    def forward_without_preconditionFusedSilu(self, x, cache):
        """linear.py:291-376 — gate_proj: no re-quantisation, consumes what up_proj left in the cache; SiLU epilogue."""
        inputs = x.reshape(-1, x.shape[-1])
        self._require_cuda(inputs, self.q_weight)
        M = inputs.shape[0]
        if self.forward_without_precondition_len != cache.ind.shape[0]:
            if cache.ind.shape[0]:
                ind = cache.new_ind
                n0 = self._n_ind
                self._gather_weight_columns(ind.contiguous(), n0)
                self.ind = cache.ind
                self.forward_without_precondition_len = self._n_ind
        if self.bit == 4 and self._n_ind == 0:
            raise RuntimeError("int4 mod should have outliers !")
        y = torch.empty((M, self.out_features), dtype=torch.float16, device=inputs.device)
        ao, ld = self._cached_act_outliers(cache, M)
        self._launch(cache, M, y, skip_prologue=True, act=ACT_SILU, q_x=cache.q_xcache, act_outliers=ao, ld_ao=ld)
        return y.reshape(cache.shape)

    # ------------------------------------------------------------------ B200 extension: RMSNorm folded into phase A
    @torch.no_grad()
    def forward_norm_fused(self, x, norm_weight, eps, cache=None, norm_out=None, residual=None):
        """RMSNorm (fused/norm.py:24-33) + this Linear in ONE launch; steady state only (after discovery).
        x is the un-normed residual stream and is left untouched."""
        if cache is None:
            cache = self.cache
        if self.add_outliers:
            raise _lib.MixqError("forward_norm_fused is a steady-state path: run the discovery calls first")
        cache.shape = x.shape[:-1] + (self.out_features,)
        inputs = x.reshape(-1, x.shape[-1])
        self._require_cuda(inputs, self.q_weight)
        M = inputs.shape[0]
        y = torch.empty((M, self.out_features), dtype=torch.float16, device=inputs.device)
        self._launch(cache, M, y, x=inputs, norm_weight=norm_weight, eps=eps, norm_out=norm_out,
                     residual=None if residual is None else residual.reshape(M, self.out_features))
        cache.ind = self.ind
        cache.q_xcache = cache.q_x_buffer(M, self.in_features)
        cache.activation_outliers = cache.ao_buffer(self._n_ind)[:M, : self._n_ind]
        return y.reshape(cache.shape)

    @torch.no_grad()
    def forward_swiglu_fused(self, up, x, norm_weight=None, eps=0.0, cache=None):
        """B200 extension: self is gate_proj, `up` is up_proj.  [RMSNorm ->] quantise once -> both GEMMs -> fp16(silu(gate)) *
        fp16(up) in ONE launch: fused/norm.py:24-33 + fused/mlp.py:61-64 (up_proj, gate_proj.forward_without_precondition-
        FusedSilu, gate *= up).  Steady state only; needs M > 128 and bit 8 (MixqError otherwise: callers fall back to the
        three-launch sequence)."""
        if cache is None:
            cache = self.cache
        if self.add_outliers and up.add_outliers:
            raise _lib.MixqError("forward_swiglu_fused is a steady-state path: run the discovery calls first")
        if self._n_ind != up._n_ind or self.in_features != up.in_features or self.out_features != up.out_features:
            raise _lib.MixqError("gate_proj and up_proj must share shape and outlier set")
        if self._wc_buf.shape[1] != up._wc_buf.shape[1]:
            n = max(self._wc_buf.shape[1], up._wc_buf.shape[1])
            self._reserve_wc(n)
            up._reserve_wc(n)
        cache.shape = x.shape[:-1] + (self.out_features,)
        inputs = x.reshape(-1, x.shape[-1])
        self._require_cuda(inputs, self.q_weight)
        M = inputs.shape[0]
        y = torch.empty((M, self.out_features), dtype=torch.float16, device=inputs.device)
        self._launch(cache, M, y, x=inputs, norm_weight=norm_weight, eps=eps, up=up)
        cache.ind = self.ind
        cache.q_xcache = cache.q_x_buffer(M, self.in_features)
        cache.activation_outliers = cache.ao_buffer(self._n_ind)[:M, : self._n_ind]
        return y.reshape(cache.shape)

    def forward_swiglu_quantized(self, up, M, cache=None):
        """forward_swiglu_fused when a producer kernel (the tensor-parallel exchange, mixq_exchange_finish_poll_quant) has
        already normalised and quantised the rows: consumes cache.q_xcache / x_scale / activation_outliers like
        forward_quantized, both GEMMs + SiLU + gate * up in one launch.  Steady state only."""
        if cache is None:
            cache = self.cache
        if self.add_outliers and up.add_outliers:
            raise _lib.MixqError("forward_swiglu_quantized is a steady-state path: run the discovery calls first")
        if self._n_ind != up._n_ind or self.in_features != up.in_features or self.out_features != up.out_features:
            raise _lib.MixqError("gate_proj and up_proj must share shape and outlier set")
        if self._wc_buf.shape[1] != up._wc_buf.shape[1]:
            n = max(self._wc_buf.shape[1], up._wc_buf.shape[1])
            self._reserve_wc(n)
            up._reserve_wc(n)
        self._require_cuda(cache.q_xcache, self.q_weight)
        y = torch.empty((M, self.out_features), dtype=torch.float16, device=self.q_weight.device)
        cache.shape = (M, self.out_features)
        ao, ld = self._cached_act_outliers(cache, M)
        self._launch(cache, M, y, skip_prologue=True, q_x=cache.q_xcache, act_outliers=ao, ld_ao=ld, up=up)
        return y

    def extra_repr(self):
        return f"in={self.in_features}, out={self.out_features}, bit={self.bit}, outliers={self._n_ind}"
