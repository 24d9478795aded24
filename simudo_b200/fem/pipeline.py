"""Host-buffer assembly with copy / compute overlap.

``PipelinedAssembler`` is the call for users whose state lives in host memory
(the reference keeps everything in host PETSc vectors, ``newton_solver.py:
61-77``): every ``submit`` copies that step's inputs from pinned host memory,
assembles (A, b) and copies b back.  Inputs and the residual are double
buffered and the three activities run on their own CUDA streams, so step i's
device->host read, step i+1's kernels and step i+2's host->device copy
overlap; ``vals`` (the 8.9 GB Jacobian at 1M cells) stays on the device for the
linear solve.
"""
import torch

__all__ = ['PipelinedAssembler']


class PipelinedAssembler(object):
    def __init__(self, plan, host_inputs, bc_values=None, natbc_values=None, depth=2,
                 pre_assemble=None, resident=None):
        """``host_inputs``: {name: pinned host tensor} with names among ``u``,
        ``base_w``, ``thmeq_phi``, ``phi_opt`` (defines shapes): what the caller changes
        between steps and is copied every step.  ``resident``: {name: device tensor} of the
        inputs that do not change per Newton step (``thermal_equilibrium_phi``, the optical
        flux between optical updates) and stay on the device.  ``pre_assemble(d)``:
        optional callable run on the compute stream before the assembly (e.g. the
        ghost exchange of a multi-GPU strip decomposition)."""
        self.plan = plan
        dev = plan.device
        self.depth = depth
        self.keys = list(host_inputs)
        self.dev_in = [{k: torch.empty_like(v, device=dev) for k, v in host_inputs.items()}
                       for _ in range(depth)]
        self.dev_b = [plan.new_vector() for _ in range(depth)]
        self.vals = plan.new_values()
        self.bc_values, self.natbc_values = bc_values, natbc_values
        self.pre_assemble = pre_assemble
        self.resident = dict(resident or {})
        self.s_h2d = torch.cuda.Stream(dev)
        self.s_cmp = torch.cuda.Stream(dev)
        self.s_d2h = torch.cuda.Stream(dev)
        mk = lambda: [torch.cuda.Event() for _ in range(depth)]
        self.ev_h2d, self.ev_cmp, self.ev_d2h = mk(), mk(), mk()
        self.step = 0

    def submit(self, host_inputs, host_b):
        """Enqueue one step; returns immediately.  ``host_b`` (pinned) holds the
        residual once ``wait()`` (or a later ``submit`` reusing its slot) returns."""
        s = self.step % self.depth
        first_use = self.step < self.depth
        d = self.dev_in[s]
        with torch.cuda.stream(self.s_h2d):
            if not first_use:
                self.s_h2d.wait_event(self.ev_cmp[s])      # inputs of step - depth consumed
            for k in self.keys:
                d[k].copy_(host_inputs[k], non_blocking=True)
            self.ev_h2d[s].record(self.s_h2d)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(self.ev_h2d[s])
            if not first_use:
                self.s_cmp.wait_event(self.ev_d2h[s])      # b slot read back
            if self.pre_assemble is not None:
                self.pre_assemble(d)
            d = dict(self.resident, **d)
            self.plan.assemble(d['u'], d.get('base_w'), d.get('thmeq_phi'), d.get('phi_opt'),
                               self.bc_values, self.natbc_values, self.vals, self.dev_b[s],
                               stream=self.s_cmp)
            self.ev_cmp[s].record(self.s_cmp)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(self.ev_cmp[s])
            host_b.copy_(self.dev_b[s], non_blocking=True)
            self.ev_d2h[s].record(self.s_d2h)
        self.step += 1

    def join(self, stream=None):
        """Make ``stream`` (default: current) wait for everything submitted."""
        stream = stream or torch.cuda.current_stream(self.plan.device)
        for s in range(min(self.step, self.depth)):
            stream.wait_event(self.ev_d2h[s])
            stream.wait_event(self.ev_cmp[s])

    def wait(self):
        self.s_d2h.synchronize()
        self.s_cmp.synchronize()
