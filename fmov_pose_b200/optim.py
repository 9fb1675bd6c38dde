"""FlatAdam — the optimiser tail of the train iteration on two kernels of this library (csrc/optim.cu).

The reference keeps one `torch.optim.Adam` over the networks (exp_runner.py:264-269) and one per pose MLP (:258-262) and,
per iteration, steps the network optimiser and ONLY the pose optimisers of the frames it rendered (:785-816).  Here every
optimiser is a *group*: [networks | pose MLP 0 | pose MLP 1 | ...]; a step gathers the gradients of this rank's active
groups into one flat fp32 buffer (plus one flag per group), all-reduces that buffer when the ray batch is sharded (the
union of the ranks' active groups is stepped — and nothing else, so pose MLPs no rank rendered keep their parameters,
moments and step counters), and applies Adam to every tensor of every active group in one launch.  Learning rates,
step counters and flags are device memory: a CUDA-graph replay follows the LR schedule.

Same update rule as `torch.optim.Adam(betas=(0.9, 0.999), eps=1e-8)` without weight decay / amsgrad (what the reference
constructs); parity over 30 steps in tests/test_gpu_optim.py."""
import ctypes

import torch

from . import _lib as L


class FlatAdam:
    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-8, process_group=None):
        """groups: list of dict(params=[...], lr=float) — group 0 = the networks, one group per pose MLP after it"""
        self.param_groups = []
        self.process_group = process_group
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        params, group_of = [], []
        for gi, g in enumerate(groups):
            ps = [p for p in g["params"]]
            for p in ps:
                assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous(), "FlatAdam takes contiguous fp32 CUDA parameters"
                params.append(p)
                group_of.append(gi)
        assert params, "no parameters"
        self.params, self.group_of = params, group_of
        self.n_groups = len(groups)
        dev = params[0].device
        self.device = dev
        chunk = int(L.lib().fmov_adam_chunk())
        offs, numels, o = [], [], 0
        for p in params:
            offs.append(o)
            numels.append(p.numel())
            o += (p.numel() + 3) // 4 * 4            # 16-byte aligned slices
        self.n_total = o
        self.offs, self.numels = offs, numels
        self.index_of = {id(p): i for i, p in enumerate(params)}
        # flat buffers: gradients (+ group flags), first and second moments
        self.G = torch.zeros(o + self.n_groups, dtype=torch.float32, device=dev)
        self.M = torch.zeros(o, dtype=torch.float32, device=dev)
        self.V = torch.zeros(o, dtype=torch.float32, device=dev)
        self.lr = torch.tensor([float(g["lr"]) for g in groups], dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(self.n_groups, dtype=torch.float32, device=dev)
        self._done = torch.zeros(1, dtype=torch.int32, device=dev)
        self._flag_cache = {}
        # device tables of fmov_adam_step
        ct, ce = [], []
        for i, n in enumerate(numels):
            for e0 in range(0, n, chunk):
                ct.append(i)
                ce.append(e0)
        self.n_chunks = len(ct)
        i64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        self._t_param = i64([p.data_ptr() for p in params])
        self._t_off, self._t_numel, self._t_group = i64(offs), i32(numels), i32(group_of)
        self._t_ct, self._t_ce = i32(ct), i32(ce)
        for gi, g in enumerate(groups):
            # param_groups[i]["lr"] is a one-element view of the device LR vector: `grp["lr"].fill_(v)` (TrainStep._write_lr)
            self.param_groups.append(dict(params=list(g["params"]), lr=self.lr[gi:gi + 1]))

    # ---- torch.optim-like surface -----------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def state_of(self, p):
        """views of the flat state for one parameter (tests / checkpoints)"""
        i = self.index_of[id(p)]
        o, n = self.offs[i], self.numels[i]
        return dict(step=self.step_count[self.group_of[i]], exp_avg=self.M[o:o + n].view_as(p),
                    exp_avg_sq=self.V[o:o + n].view_as(p))

    def state_dict(self):
        return dict(M=self.M.clone(), V=self.V.clone(), step=self.step_count.clone(), lr=self.lr.clone())

    def load_state_dict(self, sd):
        self.M.copy_(sd["M"])
        self.V.copy_(sd["V"])
        self.step_count.copy_(sd["step"])
        self.lr.copy_(sd["lr"])

    def flags_for(self, active):
        """device flag vector [n_groups] of an active set (only needed beyond 64 groups, where the launch-constant bit mask
        does not reach).  Cached per set; creating one is a host-to-device copy, so a CUDA-graph capture must find it in
        the cache (GraphedTrainStep calls this before capturing)."""
        key = tuple(sorted(int(g) for g in active))
        f = self._flag_cache.get(key)
        if f is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FlatAdam: the flag vector of this active set must exist before the capture "
                                   "(call optimizer.flags_for(active_groups) first)")
            host = torch.zeros(self.n_groups, dtype=torch.float32)
            host[list(key)] = 1.0
            f = host.to(self.device)
            self._flag_cache[key] = f
        return f

    # ---- the step ----------------------------------------------------------------------------------------------
    def gather(self, active_groups=None):
        """gradients of this rank's active groups -> G (zeroed first), flags of those groups set.  `active_groups`:
        iterable of group indices, None = every group that holds a gradient."""
        have = [(i, p.grad) for i, p in enumerate(self.params) if p.grad is not None]
        if active_groups is None:
            active = sorted({self.group_of[i] for i, _ in have})
        else:
            active = sorted(set(int(g) for g in active_groups))
            have = [(i, g) for i, g in have if self.group_of[i] in active]
        assert all(0 <= g < self.n_groups for g in active)
        mask, flags = 0, None
        if self.n_groups <= 64:
            for g in active:
                mask |= 1 << g
        else:       # long sequences (one pose MLP per frame): the flags come from a cached device vector per active set
            flags = self.flags_for(active)
        self.G.zero_()
        lib = L.lib()
        for s in range(0, max(len(have), 1), 160):
            part = have[s:s + 160]
            grads = [L.f32c(g) for _, g in part]
            n = len(part)
            src = (ctypes.c_void_p * max(n, 1))(*[g.data_ptr() for g in grads])
            off = (ctypes.c_longlong * max(n, 1))(*[self.offs[i] for i, _ in part])
            num = (ctypes.c_int * max(n, 1))(*[self.numels[i] for i, _ in part])
            L.check(lib.fmov_grad_gather(n, src, off, num, self.n_groups, ctypes.c_ulonglong(mask), L.ptr(flags),
                                         L.ptr(self.G), L.c_ll(self.n_total), L.stream()), "fmov_grad_gather")
        return active

    def step(self, active_groups=None, grad_scale=1.0):
        self.gather(active_groups)
        if self.process_group is not None:
            torch.distributed.all_reduce(self.G, group=self.process_group)
        L.check(L.lib().fmov_adam_step(L.ptr(self._t_param), L.ptr(self._t_off), L.ptr(self._t_numel), L.ptr(self._t_group),
                                       L.ptr(self._t_ct), L.ptr(self._t_ce), self.n_chunks, self.n_groups, L.ptr(self.G),
                                       L.c_ll(self.n_total), L.ptr(self.M), L.ptr(self.V), L.ptr(self.lr),
                                       L.ptr(self.step_count), L.c_float(self.betas[0]), L.c_float(self.betas[1]),
                                       L.c_float(self.eps), L.c_float(grad_scale), L.ptr(self._done), L.stream()),
                "fmov_adam_step")
