"""GaussianDiffusion / SpacedDiffusion — sampling half of MuseDiffusion/models/diffusion.py, same names and
signatures, with the per-step arithmetic executed by the fused sm_100a kernels of the C-ABI library.

What is mirrored (reference file:line):
  get_named_beta_schedule / betas_for_alpha_bar(_left)   diffusion.py:22-118
  GaussianDiffusion.__init__ tables                       :136-185   (float64 numpy, cast to fp32 at upload as :914 does)
  q_sample                                                :229-255
  q_posterior_mean_variance / p_mean_variance / p_sample  :257-404
  p_sample_loop(_progressive)                             :406-540
  ddim_sample / ddim_sample_loop(_progressive)            :701-757, :797-901
  space_timesteps / SpacedDiffusion / _WrappedModel       :920-1032
Training losses (:187-192, :542-699) are out of scope and raise NotImplementedError.

Noise.  The reference draws `torch.randn_like` every step (plus a host-synchronising rejection loop for top_p).
Here noise is generated inside the posterior kernel by a counter-based Philox stream keyed by
(torch.initial_seed(), call counter, global element index), so results do not depend on how a batch is sharded.
Set `diffusion.noise_source = fn(shape, kind) -> Tensor` (kind in {"randn", "truncated"}) to inject external noise
instead — the parity tests do this to feed the oracle's exact stream."""
import math

import numpy as np
import torch

from . import _lib, ops
from .rounding import rounding_weight_of


# ------------------------------------------------------------------------------------------------ schedules
def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """diffusion.py:101-118."""
    T = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / T) / alpha_bar(i / T), max_beta) for i in range(T)])


def betas_for_alpha_bar_left(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """diffusion.py:80-98: like betas_for_alpha_bar with an extra leading beta and one step fewer."""
    T = num_diffusion_timesteps
    head = [min(1 - alpha_bar(0), max_beta)]
    return np.array(head + [min(1 - alpha_bar((i + 1) / T) / alpha_bar(i / T), max_beta) for i in range(T - 1)])


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """diffusion.py:22-77."""
    T = num_diffusion_timesteps
    scale = 1000 / T
    if schedule_name == "linear":
        return np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(T, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    if schedule_name == "sqrt":
        return betas_for_alpha_bar(T, lambda t: 1 - np.sqrt(t + 0.0001))
    if schedule_name == "trunc_cos":
        return betas_for_alpha_bar_left(T, lambda t: np.cos((t + 0.1) / 1.1 * np.pi / 2) ** 2)
    if schedule_name == "trunc_lin":
        return np.linspace(scale * 0.0001 + 0.01, scale * 0.02 + 0.01, T, dtype=np.float64)
    if schedule_name == "pw_lin":
        lo, mid, hi = scale * 0.0001 + 0.01, scale * 0.0001, scale * 0.02
        return np.concatenate([np.linspace(lo, mid, 10, dtype=np.float64), np.linspace(mid, hi, T - 10, dtype=np.float64)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    """diffusion.py:920-969."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    per, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, count in enumerate(section_counts):
        size = per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            steps.append(start + round(pos))
            pos += stride
        start += size
    return set(steps)


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """diffusion.py:904-917 (API parity; the kernels read the same fp32 values from constant memory)."""
    res = torch.from_numpy(np.asarray(arr)).to(device=timesteps.device)[timesteps].float()
    while res.dim() < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


# ------------------------------------------------------------------------------------------------ diffusion
_capture_state = {}          # device index -> [most recent graph, capture stream], kept for the life of the process


def _capture(dev, fn):
    """Record `fn`'s launches into a new CUDA graph.  Plain capture_begin / capture_end on a side stream instead of the
    `torch.cuda.graph` context: the context empties the caching allocator (device and pinned host) on entry, which at
    the bench size means returning and re-acquiring ~10 GB of step activations on EVERY sampling call (60-230 ms per
    call, measured with tools/e2e_overhead.py).  Every capture on a device allocates from the memory pool of the previous
    one (which is kept alive until the new graph exists, so the pool never dies): the step's temporaries are free blocks
    of that pool again by the time a capture ends, and the next call's capture takes the same blocks.  Only the newest
    graph of a device is ever replayed."""
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    state = _capture_state.setdefault(index, [None, torch.cuda.Stream(device=dev)])
    previous, side = state
    graph = torch.cuda.CUDAGraph()
    main = torch.cuda.current_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        graph.capture_begin(pool=previous.pool() if previous is not None else torch.cuda.graph_pool_handle())
        try:
            fn()
        finally:
            graph.capture_end()
    main.wait_stream(side)
    state[0] = graph
    return graph


class GaussianDiffusion:
    """Sampling utilities with the reference's attribute and method names (diffusion.py:121-901)."""

    noise_source = None     # optional callable(shape, kind) -> CUDA float tensor; None = in-kernel Philox
    rounding_trace = None   # optional list: every fused rounding call appends (ids int32 [M], top-2 margin fp32 [M]) clones
    use_cuda_graph = True   # loops of >= GRAPH_MIN_STEPS steps replay ONE captured CUDA graph of a reverse step (fast path only)
    GRAPH_MIN_STEPS = 4
    seq_offset = 0          # global index of this rank's first sequence (keeps Philox noise shard-invariant)

    def __init__(self, *, betas, predict_xstart, rescale_timesteps=False, training_mode="s2s"):
        self.rescale_timesteps = rescale_timesteps
        self.predict_xstart = predict_xstart
        self.training_mode = training_mode
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        # fixed-large model variance used by p_mean_variance (diffusion.py:313-317)
        self.model_variance = np.append(self.posterior_variance[1], self.betas[1:])
        self.model_log_variance = np.log(self.model_variance)
        self._noise_calls = 0
        self._sched_token = object()

    # -------------------------------------------------------------------------------------- plumbing
    def _upload_schedule(self):
        """fp32 coefficient tables -> __constant__ memory (once; re-uploaded only if another schedule displaced them)."""
        if ops.current_schedule_key() is not self._sched_token:
            ops.set_schedule({n: getattr(self, n) for n in ops.TABLE_ORDER}, key=self._sched_token)

    def _next_counter(self):
        self._noise_calls += 1
        return self._noise_calls

    @staticmethod
    def _seed():
        return int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF

    def _external_noise(self, shape, kind, device):
        if self.noise_source is None:
            return None
        n = self.noise_source(tuple(shape), kind)
        if not isinstance(n, torch.Tensor):
            n = torch.from_numpy(np.ascontiguousarray(n))
        return n.to(device=device, dtype=torch.float32)

    def _scale_timesteps(self, t):
        """diffusion.py:207-210."""
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def training_losses(self, *a, **k):
        raise NotImplementedError("training is outside the scope of musediffusion_b200 (sampling hot path only)")

    # -------------------------------------------------------------------------------------- forward noising
    def q_sample(self, x_start, t, noise=None, mask=None):
        """diffusion.py:229-255.  x_start [B, ...]; t holds one index per sequence ([B] or [B, 1], as
        run/sample.py:196-197 passes it); mask (already broadcastable to x_start once unsqueezed) keeps x_start
        where it is 0."""
        self._upload_schedule()
        shape = x_start.shape
        B = shape[0]
        x3, m3 = x_start, mask
        if x_start.dim() == 4 and shape[-1] == 1:          # run/sample.py:197 passes [B, L, D, 1]
            x3 = x_start.squeeze(-1)
        elif x_start.dim() != 3:
            x3 = x_start.reshape(B, 1, -1)
            if mask is not None:
                m3 = torch.broadcast_to(mask.unsqueeze(-1), shape).reshape(B, 1, -1)
        if noise is None:
            noise = self._external_noise(shape, "randn", x_start.device)
        if noise is not None:
            noise = noise.reshape(x3.shape)
        out = ops.q_sample(x3, t.reshape(-1), noise=noise, seed=self._seed(), step_counter=self._next_counter(),
                           seq_offset=self.seq_offset, mask=m3)
        return out.view(shape)

    def _predict_xstart_from_eps(self, x_t, t, eps):
        """diffusion.py:194-199."""
        self._upload_schedule()
        return ops.xstart_from_eps(x_t, eps, t)

    # -------------------------------------------------------------------------------------- one reverse step
    def _model_timesteps(self, t):
        """value fed to the denoiser for schedule index t (overridden by SpacedDiffusion)."""
        return self._scale_timesteps(t)

    def _call_model(self, model, x, t, model_kwargs):
        """diffusion.py:309: `model(x, self._scale_timesteps(t), model_kwargs=model_kwargs)` — the kwarg is passed
        whole and the reference model drops it (network.py:131), which is preserved for arbitrary callables."""
        return model(x, self._model_timesteps(t), model_kwargs=model_kwargs if model_kwargs is not None else {})

    def _step(self, mode, model, x, t, clip_denoised, denoised_fn, model_kwargs, top_p, mask, x_start, eta,
              want_aux=True, model_output=None, out=None, out_bf16=None, uniform_t=False, step_counter=None):
        """Shared body of p_sample / ddim_sample: model call, rounding, fused posterior kernel."""
        self._upload_schedule()
        B = x.shape[0]
        assert t.numel() in (1, B)
        if model_output is None:
            model_output = self._call_model(model, x, t, model_kwargs)
        if not self.predict_xstart:
            model_output = ops.xstart_from_eps(x, model_output, t)
        E = rounding_weight_of(denoised_fn)
        idx = pred = None
        if E is not None:
            # fused distance contraction + row argmin (rounding.py:21-28); tcgen05 split-bf16 kernel when the embedding
            # width allows it, fp32 CUDA-core kernel otherwise
            trace = self.rounding_trace
            if E.shape[1] % 64 == 0:
                idx = ops.round_argmin_tc(model_output, ops.split_embedding(E), want_margin=trace is not None)
            else:
                idx = ops.round_argmin(model_output, E, want_margin=trace is not None)
            if trace is not None:
                idx, margin = idx
                trace.append((idx.clone(), margin))
        elif denoised_fn is not None:
            pred = denoised_fn(model_output, t if t.numel() == B else t.expand(B))   # arbitrary user callable
        else:
            pred = model_output
        if mode == _lib.STEP_DDPM:
            kind, tp = ("truncated", top_p) if (top_p is not None and top_p > 0) else ("randn", 0.0)
        else:
            kind, tp = "randn", 0.0                                  # ddim_sample ignores top_p (diffusion.py:738)
        noise = self._external_noise(x.shape, kind, x.device)
        pred_out = torch.empty_like(x, dtype=torch.float32) if want_aux else None
        mean_out = torch.empty_like(x, dtype=torch.float32) if want_aux else None
        sample = ops.posterior_step(
            x, t, mode, idx=idx, pred=pred, E=E, noise=noise, seed=self._seed(),
            step_counter=self._next_counter() if step_counter is None else step_counter, seq_offset=self.seq_offset,
            mask=mask, x_start=x_start, eta=eta, clip=clip_denoised, top_p=tp, out=out, out_bf16=out_bf16,
            pred_out=pred_out, mean_out=mean_out)
        return sample, pred_out, mean_out

    def q_posterior_mean_variance(self, x_start, x_t, t):
        """diffusion.py:257-278."""
        self._upload_schedule()
        mean = torch.empty_like(x_t, dtype=torch.float32)
        zeros = torch.zeros_like(mean)
        ops.posterior_step(x_t, t, _lib.STEP_DDPM, pred=x_start, noise=zeros, clip=False, mean_out=mean)
        return (mean, _extract_into_tensor(self.posterior_variance, t, x_t.shape),
                _extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape))

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """diffusion.py:280-347: dict(mean, variance, log_variance, pred_xstart)."""
        zeros = torch.zeros_like(x, dtype=torch.float32)
        saved, self.noise_source = self.noise_source, (lambda shape, kind: zeros)
        try:
            _, pred, mean = self._step(_lib.STEP_DDPM, model, x, t, clip_denoised, denoised_fn, model_kwargs, None,
                                       None, None, 0.0)
        finally:
            self.noise_source = saved
        return {"mean": mean, "variance": _extract_into_tensor(self.model_variance, t, x.shape),
                "log_variance": _extract_into_tensor(self.model_log_variance, t, x.shape), "pred_xstart": pred}

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, top_p=None, mask=None,
                 x_start=None):
        """diffusion.py:349-404."""
        sample, pred, mean = self._step(_lib.STEP_DDPM, model, x, t, clip_denoised, denoised_fn, model_kwargs, top_p,
                                        mask, x_start, 0.0)
        out = {"mean": mean, "variance": _extract_into_tensor(self.model_variance, t, x.shape),
               "log_variance": _extract_into_tensor(self.model_log_variance, t, x.shape), "pred_xstart": pred}
        return {"sample": sample, "pred_xstart": pred, "greedy_mean": mean, "out": out}

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, eta=0.0,
                    langevin_fn=None, mask=None, x_start=None):
        """diffusion.py:701-757.  `langevin_fn(sample, mean_pred, sigma, alpha_bar_prev[t[0]], t, x)` (:748-750) is applied
        to the un-masked sample; the mask select that follows it (:752-755) is md_q_sample's `mask == 0 ? x_start : sample`."""
        if not langevin_fn:
            sample, pred, _ = self._step(_lib.STEP_DDIM, model, x, t, clip_denoised, denoised_fn, model_kwargs, None, mask,
                                         x_start, eta)
            return {"sample": sample, "pred_xstart": pred}
        sample, pred, mean_pred = self._step(_lib.STEP_DDIM, model, x, t, clip_denoised, denoised_fn, model_kwargs, None, None,
                                             None, eta)
        tt = t.reshape(-1).long()
        ab, abp = self.alphas_cumprod, self.alphas_cumprod_prev
        sigma_tab = eta * np.sqrt((1 - abp) / (1 - ab)) * np.sqrt(1 - ab / abp)
        sigma = _extract_into_tensor(sigma_tab, tt if tt.numel() == x.shape[0] else tt.expand(x.shape[0]), x.shape)
        sample = langevin_fn(sample, mean_pred, sigma, self.alphas_cumprod_prev[int(tt[0])], t, x)
        if mask is not None:
            sample = ops.q_sample(sample.new_empty(0) if x_start is None else x_start.to(torch.float32).contiguous(), None,
                                  noise=sample.to(torch.float32).contiguous(), mask=mask)
        return {"sample": sample, "pred_xstart": pred}

    # -------------------------------------------------------------------------------------- loops
    def _fast_model(self, model):
        """our CUDA denoiser behind `model` (possibly wrapped), or None for arbitrary callables."""
        from .network import TransformerNetModel
        inner = model.model if isinstance(model, _WrappedModel) else model
        return inner if isinstance(inner, TransformerNetModel) else None

    def _loop(self, mode, model, shape, noise, clip_denoised, denoised_fn, model_kwargs, device, progress, top_p,
              clamp_step, clamp_first, mask, x_start, eta, indices, want_aux):
        """Generator over steps.  Fast path (our denoiser): no per-step allocation, x ping-pongs between two
        buffers, the posterior kernel also emits the bf16 operand of the next step's first GEMM, one time-embedding
        row per step (all sequences share t inside the loops, diffusion.py:516,887)."""
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        if noise is not None:
            x = noise
        else:
            x = self._external_noise(shape, "randn", device)
            if x is None:
                x = ops.fill_normal(tuple(shape), device, seed=self._seed(), step_counter=self._next_counter(),
                                    elem_offset=self.seq_offset * int(np.prod(shape[1:])))
        fast = self._fast_model(model)
        B = shape[0]
        if len(indices) == 0:
            return
        if mask is not None:
            # convert the (usually int64, stride-0 expanded) mask ONCE so the per-step kernel call allocates nothing
            m = torch.broadcast_to(mask if mask.dim() == len(shape) else mask.unsqueeze(-1), tuple(shape))
            if m.stride(-1) == 0:
                mask = m[..., 0].to(torch.int32).contiguous().unsqueeze(-1).expand(tuple(shape))
            else:
                mask = m.to(torch.int32).contiguous()
        if x_start is not None:
            x_start = x_start.to(torch.float32).contiguous()
        x = x.to(torch.float32).contiguous()
        idx_list = list(indices)
        t_idx = torch.tensor(idx_list, dtype=torch.int32, device=device)
        t_model = self._model_timesteps(t_idx.long()).float()
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(idx_list)
        else:
            indices = idx_list

        def rounds_at(i):
            """clamp gating, diffusion.py:517-526 (the DDIM loop always rounds, :889-899); like the reference, the DDPM loop
            raises TypeError when clamp_step is left at its default None"""
            if mode != _lib.STEP_DDPM:
                return True
            return (i >= clamp_step) if clamp_first else not (i > clamp_step)

        E = rounding_weight_of(denoised_fn)
        graph_ok = (fast is not None and self.use_cuda_graph and len(idx_list) >= self.GRAPH_MIN_STEPS and not want_aux
                    and self.noise_source is None and self.rounding_trace is None and ops._PROFILE[0] is None
                    and E is not None and E.shape[1] % 64 == 0 and all(rounds_at(i) for i in idx_list)
                    and x.dim() == 3)
        if graph_ok:
            yield from self._graph_loop(mode, fast, x, E, t_idx, t_model, indices, clip_denoised, top_p, mask, x_start, eta)
            return
        bufs = [torch.empty(tuple(shape), dtype=torch.float32, device=device) for _ in range(2)] if not want_aux else None
        xb = [torch.empty(tuple(shape), dtype=torch.bfloat16, device=device) for _ in range(2)] if fast is not None else None
        x_bf16 = None
        mo_buf = torch.empty(tuple(shape), dtype=torch.float32, device=device) if fast is not None else None
        for k, i in enumerate(indices):
            fn = denoised_fn if rounds_at(i) else None
            t1 = t_idx[k:k + 1]
            model_output = None
            if fast is not None:
                model_output = fast.denoise(x, t_model[k:k + 1], x_bf16=x_bf16, uniform_t=True, out=mo_buf)
            t_arg = t1 if fast is not None else t_idx[k:k + 1].long().expand(B)
            sample, pred, mean = self._step(
                mode, model, x, t_arg, clip_denoised, fn, model_kwargs, top_p, mask, x_start, eta, want_aux=want_aux,
                model_output=model_output, out=None if bufs is None else bufs[k & 1],
                out_bf16=None if xb is None else xb[k & 1])
            x = sample
            x_bf16 = None if xb is None else xb[k & 1]
            yield sample, pred, mean, t_arg

    def _graph_loop(self, mode, fast, x, E, t_idx, t_model, indices, clip_denoised, top_p, mask, x_start, eta):
        """The same chain with ONE reverse step captured as a CUDA graph and replayed.
        What changes from step to step lives in device memory: `md_step_advance` (first node of the graph) moves a cursor
        over the index tables and publishes the schedule index, the denoiser's timestep value and the Philox counter of
        the step; the posterior kernel reads the counter from there, so the noise is bit-identical to the eager loop.
        x_t is updated in place (the posterior kernel is elementwise) and its bf16 copy feeds the next step's first GEMM.
        The first step runs eagerly (one-time set-up: shared-memory attributes, weight pack, tensor maps), the graph is
        captured from the second one."""
        dev = x.device
        n = t_idx.numel()
        B = x.shape[0]
        self._upload_schedule()
        x = x.clone()                                   # updated in place: never the caller's tensor
        xb = ops.cast_bf16(x)
        # x0-predicting model: the last Linear emits the [hi | lo] bf16 split that the rounding contraction consumes, the fp32
        # model output never exists; an eps-predicting model needs it in fp32 for x0 = sr x_t - srm1 eps first
        use_split = self.predict_xstart and fast.weight_pack().has_down
        mo = torch.empty_like(x) if not use_split else None
        mo_split = torch.empty(x.shape[:-1] + (2 * x.shape[-1],), dtype=torch.bfloat16, device=dev) if use_split else None
        cursor = torch.zeros(1, dtype=torch.int32, device=dev)
        t_cur = torch.zeros(1, dtype=torch.int32, device=dev)
        tm_cur = torch.zeros(1, dtype=torch.float32, device=dev)
        ctr_cur = torch.zeros(1, dtype=torch.int64, device=dev)
        ctr_base = self._noise_calls + 1                # the value _next_counter() hands to the first step of an eager loop
        seed = self._seed()
        se = ops.split_embedding(E)
        # clip_denoised after rounding = gathering rows of clamp(E, -1, 1): built once per embedding matrix, not per token and step
        E_step, clip_step = (se.E_clamped, 2) if clip_denoised else (E, 0)
        idx = torch.empty((x.shape[0] * x.shape[1],), dtype=torch.int32, device=dev)
        if mode == _lib.STEP_DDPM:
            tp = top_p if (top_p is not None and top_p > 0) else 0.0
        else:
            tp = 0.0                                    # ddim_sample ignores top_p (diffusion.py:738)

        def one_step():
            ops.step_advance(cursor, t_idx, t_model, t_cur, tm_cur, ctr_cur, ctr_base)
            if use_split:
                fast.denoise(x, tm_cur, x_bf16=xb, uniform_t=True, split_out=mo_split)
                ops.round_argmin_tc(None, se, out=idx, presplit=mo_split)
            else:
                out = fast.denoise(x, tm_cur, x_bf16=xb, uniform_t=True, out=mo)
                if not self.predict_xstart:
                    out = ops.xstart_from_eps(x, out, t_cur)
                ops.round_argmin_tc(out, se, out=idx)
            ops.posterior_step(x, t_cur, mode, idx=idx, E=E_step, seed=seed, seq_offset=self.seq_offset, mask=mask, x_start=x_start,
                               eta=eta, clip=clip_step, top_p=tp, out=x, out_bf16=xb, step_counter_dev=ctr_cur)

        graph, nodes = None, 0
        for k, i in enumerate(indices):
            if k == 0:
                one_step()
            else:
                if graph is None:
                    before = ops.launch_count()
                    graph = _capture(dev, one_step)
                    nodes = ops.launch_count() - before          # kernels recorded by the capture (nothing ran yet)
                    ops.add_launches(-nodes)
                graph.replay()
                ops.add_launches(nodes)
            self._noise_calls += 1
            yield x, None, None, t_idx[k:k + 1]

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                      device=None, progress=False, top_p=None, clamp_step=None, clamp_first=None, mask=None,
                      x_start=None, gap=1, eta=0.0, t_enc=None, only_last=False):
        """diffusion.py:406-473.  Returns a list of samples (one element when only_last)."""
        indices = list(range(self.num_timesteps))[::-1][slice(t_enc)]
        final, last = [], None
        for last in self._loop(_lib.STEP_DDPM, model, shape, noise, clip_denoised, denoised_fn, model_kwargs, device,
                               progress, top_p, clamp_step, clamp_first, mask, x_start, eta, indices, want_aux=False):
            if not only_last:
                final.append(last[0].clone())
        if only_last:
            if last is None:
                return []
            final.append(last[0])
        return final

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                  model_kwargs=None, device=None, progress=False, top_p=None, clamp_step=None,
                                  clamp_first=None, mask=None, x_start=None, eta=0.0, t_enc=None):
        """diffusion.py:475-540: generator of p_sample dicts."""
        indices = list(range(self.num_timesteps))[::-1][slice(t_enc)]
        for sample, pred, mean, t in self._loop(_lib.STEP_DDPM, model, shape, noise, clip_denoised, denoised_fn,
                                                model_kwargs, device, progress, top_p, clamp_step, clamp_first, mask,
                                                x_start, eta, indices, want_aux=True):
            yield {"sample": sample, "pred_xstart": pred, "greedy_mean": mean,
                   "out": {"mean": mean, "pred_xstart": pred,
                           "variance": _extract_into_tensor(self.model_variance, t.long(), sample.shape),
                           "log_variance": _extract_into_tensor(self.model_log_variance, t.long(), sample.shape)}}

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                         device=None, progress=False, top_p=None, clamp_step=None, clamp_first=None, mask=None,
                         x_start=None, gap=1, eta=0.0, t_enc=None, only_last=False):
        """diffusion.py:797-846 (top_p / clamp_step / clamp_first accepted and ignored, like the reference)."""
        indices = list(range(self.num_timesteps))[::-1][::gap][slice(t_enc)]
        final, last = [], None
        for last in self._loop(_lib.STEP_DDIM, model, shape, noise, clip_denoised, denoised_fn, model_kwargs, device,
                               progress, None, 0, True, mask, x_start, eta, indices, want_aux=False):
            if not only_last:
                final.append(last[0].clone())
        if only_last:
            if last is None:
                return []
            final.append(last[0])
        return final

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0, langevin_fn=None,
                                     mask=None, x_start=None, gap=1, t_enc=None):
        """diffusion.py:848-901 (`langevin_fn` is accepted and, as in the reference, never handed to ddim_sample :866-877)."""
        indices = list(range(self.num_timesteps))[::-1][::gap][slice(t_enc)]
        for sample, pred, _, _ in self._loop(_lib.STEP_DDIM, model, shape, noise, clip_denoised, denoised_fn,
                                             model_kwargs, device, progress, None, 0, True, mask, x_start, eta, indices,
                                             want_aux=True):
            yield {"sample": sample, "pred_xstart": pred}


class SpacedDiffusion(GaussianDiffusion):
    """diffusion.py:972-1018: a diffusion process that keeps only `use_timesteps` of a base process."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        last, new_betas = 1.0, []
        for i, ac in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t                                   # scaling is done by the wrapped model (diffusion.py:1015-1017)

    def _model_timesteps(self, t):
        """_WrappedModel.__call__ (diffusion.py:1027-1032) without the per-step H2D copy of the map."""
        m = getattr(self, "_map_dev", None)
        if m is None or m.device != t.device:
            m = torch.tensor(self.timestep_map, device=t.device, dtype=torch.long)
            self._map_dev = m
        new_ts = m[t.long()]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return new_ts

    def _call_model(self, model, x, t, model_kwargs):
        inner = model.model if isinstance(model, _WrappedModel) else model
        return inner(x, self._model_timesteps(t), model_kwargs=model_kwargs if model_kwargs is not None else {})

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)


class _WrappedModel:
    """diffusion.py:1020-1032."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps

    def __call__(self, x, ts, **kwargs):
        map_tensor = torch.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        new_ts = map_tensor[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)


def unwrap_model(model, unwrap_parallel=True):
    """diffusion.py:1035-1041."""
    if isinstance(model, _WrappedModel):
        return unwrap_model(model.model)
    if isinstance(model, (torch.nn.parallel.DistributedDataParallel, torch.nn.parallel.DataParallel)) and unwrap_parallel:
        return unwrap_model(model.module)
    return model
