"""Device-resident entry: torch tensors in HBM -> rehuel_irk_ensemble_device on torch's current stream.

torch is plumbing only (device memory + stream handle); the integration itself is the library's
CUDA kernel.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import capi


class DeviceEnsemble:
    """Pre-allocated device buffers for one ensemble shape, reusable across solves."""

    def __init__(self, problem: int, method: int, M: int, y0: torch.Tensor, params: torch.Tensor,
                 n_samples_cap: int = 0, n_hint: int = 0, trace_cap: int = 0, trace_member: int = 0,
                 order: torch.Tensor | None = None, rtol: torch.Tensor | None = None,
                 atol: torch.Tensor | None = None, newton_tol: torch.Tensor | None = None,
                 handout_order: int = capi.ORDER_INDEX, handout_keys: torch.Tensor | None = None):
        """``order``: caller-supplied hand-out order (int32 [M]).  ``handout_order``: ORDER_LOCALITY(_DESC) makes
        every ``solve`` recompute the parameter-locality order on the device first (rehuel_locality_order), from
        ``handout_keys`` ([k, M] float64: whatever the cost of a member depends on, e.g. parameters AND tolerance)
        or, by default, from the parameters."""
        assert y0.is_cuda and y0.dtype == torch.float64 and y0.is_contiguous()
        dev = y0.device
        n, p = capi.problem_dims(problem, n_hint or y0.shape[0])
        assert y0.shape[0] == n
        if p:
            assert params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()
            assert tuple(params.shape) == (p, M)
        self.problem, self.method, self.M, self.n, self.p, self.n_hint = problem, method, M, n, p, n_hint or n
        self.y0, self.params = y0, params
        self.y = torch.empty((n, M), dtype=torch.float64, device=dev)
        self.t_final = torch.empty(M, dtype=torch.float64, device=dev)
        self.err_last = torch.empty(M, dtype=torch.float64, device=dev)
        self.status = torch.empty(M, dtype=torch.int32, device=dev)
        self.counters = torch.empty((capi.NCOUNTERS, M), dtype=torch.int32, device=dev)
        self.n_samples = torch.empty(M, dtype=torch.int32, device=dev)
        self.queue = torch.zeros(1, dtype=torch.int64, device=dev)
        self.cap = n_samples_cap
        if n_samples_cap:
            self.t_samples = torch.empty((n_samples_cap, M), dtype=torch.float64, device=dev)
            self.y_samples = torch.empty((n_samples_cap, n, M), dtype=torch.float64, device=dev)
            self.err_samples = torch.empty((n_samples_cap, M), dtype=torch.float64, device=dev)
        self.trace_cap = trace_cap
        if trace_cap:
            self.trace = torch.zeros((trace_cap, 5), dtype=torch.float64, device=dev)
            self.trace_len = torch.zeros(1, dtype=torch.int32, device=dev)
        io = capi.DeviceIO()
        io.y0 = y0.data_ptr()
        io.y0_broadcast = int(y0.dim() == 1)
        io.params = params.data_ptr() if p else None
        io.y, io.t_final, io.err_last = self.y.data_ptr(), self.t_final.data_ptr(), self.err_last.data_ptr()
        io.status, io.counters = self.status.data_ptr(), self.counters.data_ptr()
        io.n_samples = self.n_samples.data_ptr() if n_samples_cap else None  # None: skip sample bookkeeping
        io.queue = self.queue.data_ptr()
        io.n_samples_cap = n_samples_cap
        if n_samples_cap:
            io.t_samples, io.y_samples, io.err_samples = (self.t_samples.data_ptr(), self.y_samples.data_ptr(),
                                                          self.err_samples.data_ptr())
        if trace_cap:
            io.trace, io.trace_cap, io.trace_member = self.trace.data_ptr(), trace_cap, trace_member
            io.trace_len = self.trace_len.data_ptr()
        if order is not None:
            assert order.is_cuda and order.dtype == torch.int32 and order.numel() == M
            self.order = order.contiguous()
            io.order = self.order.data_ptr()
        self.handout_order = capi.ORDER_INDEX
        self.keys = handout_keys if handout_keys is not None else (params if p else None)
        if handout_keys is not None:
            assert handout_keys.is_cuda and handout_keys.dtype == torch.float64 and handout_keys.is_contiguous()
            assert handout_keys.dim() == 2 and handout_keys.shape[1] == M
        if order is None and handout_order != capi.ORDER_INDEX and self.keys is not None:
            self.handout_order = handout_order
            self.order = torch.empty(M, dtype=torch.int32, device=dev)
            ws = capi.lib.rehuel_locality_order_workspace(M)
            if ws == 0:
                raise capi.RehuelError(-2, capi.last_error())
            self.order_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
            io.order = self.order.data_ptr()
        if rtol is not None:
            for v in (rtol, atol, newton_tol):
                assert v is not None and v.is_cuda and v.dtype == torch.float64 and v.numel() == M and v.is_contiguous()
            self.tols = (rtol, atol, newton_tol)
            io.rtol, io.atol, io.newton_tol = rtol.data_ptr(), atol.data_ptr(), newton_tol.data_ptr()
        self.io = io

    def reorder(self) -> None:
        """Enqueue the parameter-locality hand-out order (no-op for index order / caller-given order)."""
        if self.handout_order == capi.ORDER_INDEX:
            return
        stream = torch.cuda.current_stream(self.y.device).cuda_stream
        with torch.cuda.device(self.y.device):
            capi.check(capi.lib.rehuel_locality_order(
                self.keys.data_ptr(), int(self.keys.shape[0]), self.M, int(self.handout_order == capi.ORDER_LOCALITY_DESC),
                self.order.data_ptr(), self.order_ws.data_ptr(), self.order_ws.numel(), C.c_void_p(stream)))

    def launch(self, t0: float, t1: float, dt0: float, opts: capi.Options) -> None:
        """Enqueue the integration kernel alone on torch's current CUDA stream (asynchronous)."""
        stream = torch.cuda.current_stream(self.y.device).cuda_stream
        with torch.cuda.device(self.y.device):
            capi.check(capi.lib.rehuel_irk_ensemble_device(
                self.problem, self.method, t0, t1, dt0, self.M, self.n_hint, C.byref(opts), C.byref(self.io),
                C.c_void_p(stream)))

    def solve(self, t0: float, t1: float, dt0: float, opts: capi.Options) -> None:
        """One ensemble solve on torch's current CUDA stream (asynchronous): hand-out order, then the kernel."""
        self.reorder()
        self.launch(t0, t1, dt0, opts)

    def counter(self, name: str) -> torch.Tensor:
        return self.counters[capi.COUNTER_ROWS.index(name)]
