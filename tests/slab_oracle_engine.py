"""TEST INFRASTRUCTURE: a CPU stand-in for nuclearmpm_b200.slab.GpuSlabEngine built on the oracle port, so the
slab protocol of slab.SlabSimulation (ownership, ghost-plane reduction, migration records, count table,
re-balancing) runs under gloo with world_size 2 on a machine without a GPU."""
import numpy as np
import torch

from nuclearmpm_b200.slab import base_x
from oracle import cpu_oracle as co


class OracleSlabEngine:
    def __init__(self, x, ids, model, res, dt, E, nu, gravity, slab, capacity, device, **state):
        n, d = x.shape
        self.dim, self.res, self.n1 = d, res, res + 1
        self.args = (model, res, dt, E, nu, gravity)
        eye = np.zeros((d, d), np.float32)
        eye[0, 0] = eye[1, 1] = 1  # diag<dim>(1), Q1
        v0 = state.get("v")
        self.st = dict(x=x.astype(np.float32), v=np.zeros((n, d), np.float32) if v0 is None else v0.astype(np.float32),
                       F=np.tile(eye, (n, 1, 1)),
                       C=np.zeros((n, d, d), np.float32), Jp=np.ones(n, np.float32), mass=np.ones(n, np.float32),
                       volume=np.ones(n, np.float32), ids=ids.astype(np.uint32))
        self.range = tuple(slab)
        self.capacity = capacity
        self.rec_words = 2 * d + 2 * d * d + 4
        self.plane_words = self.n1 ** (d - 1) * 4
        self.cells = self.n1 ** d
        self.gridbuf = torch.zeros(self.cells * 4)
        self.pending = None
        self.launches = 0

    def new_buffer(self, n, dtype="float32"):
        return torch.zeros(int(n), dtype=getattr(torch, dtype))

    def _sim(self):
        s = self.st
        m, res, dt, E, nu, g = self.args
        return co.CpuSim(s["x"], m, res, dt, E, nu, g, v=s["v"], F=s["F"], Cm=s["C"], Jp=s["Jp"], mass=s["mass"],
                         volume=s["volume"])

    def p2g(self):
        if self.pending is not None:  # records received at the end of the previous step
            for k in self.st:
                self.st[k] = np.concatenate([self.st[k], self.pending[k]])
            self.pending = None
        assert len(self.st["x"]) <= self.capacity
        self.sim = self._sim()
        self.sim.phase(0)
        gv, gm = self.sim.grid()
        g4 = np.zeros((self.cells, 4), np.float32)
        g4[:, :self.dim] = gv
        g4[:, self.dim] = gm
        self.gridbuf = torch.from_numpy(g4.reshape(-1))

    def plane_view(self, x_plane, planes):
        return self.gridbuf[x_plane * self.plane_words:(x_plane + planes) * self.plane_words]

    def add_planes(self, x_plane, planes, buf):
        self.plane_view(x_plane, planes).add_(buf)

    def grid_g2p(self, send_left, send_right, cap_records, counts):
        d = self.dim
        g4 = self.gridbuf.numpy().reshape(self.cells, 4)
        self.sim.set_grid(np.ascontiguousarray(g4[:, :d]), np.ascontiguousarray(g4[:, d]))
        self.sim.phase(1)
        self.sim.phase(2)
        new = self.sim.particles()
        for k in ("x", "v", "F", "C", "Jp"):
            self.st[k] = new[k]
        bx = base_x(self.st["x"], self.res)
        left, right = bx < self.range[0], bx >= self.range[1]
        n = len(bx)
        rec = np.concatenate([self.st["x"], self.st["v"], self.st["F"].reshape(n, -1), self.st["C"].reshape(n, -1),
                              self.st["Jp"][:, None], self.st["mass"][:, None], self.st["volume"][:, None],
                              self.st["ids"].view(np.float32)[:, None]], axis=1).astype(np.float32)
        assert rec.shape[1] == self.rec_words
        nl, nr = int(left.sum()), int(right.sum())
        over = int(nl > cap_records or nr > cap_records)
        if not over:
            send_left[:nl * self.rec_words] = torch.from_numpy(rec[left].reshape(-1))
            send_right[:nr * self.rec_words] = torch.from_numpy(rec[right].reshape(-1))
        keep = ~(left | right)
        for k in self.st:
            self.st[k] = self.st[k][keep]
        counts[:] = torch.tensor([nl, nr, int(keep.sum()), over], dtype=torch.int32)

    def unpack(self, recv_left, n_left, recv_right, n_right, n_sent):
        d, W = self.dim, self.rec_words
        recs = np.concatenate([recv_left[:n_left * W].numpy().reshape(-1, W), recv_right[:n_right * W].numpy().reshape(-1, W)])
        k = len(recs)
        o = 0
        out = {}
        for name, w, shape in (("x", d, (k, d)), ("v", d, (k, d)), ("F", d * d, (k, d, d)), ("C", d * d, (k, d, d)),
                               ("Jp", 1, (k,)), ("mass", 1, (k,)), ("volume", 1, (k,))):
            out[name] = np.ascontiguousarray(recs[:, o:o + w]).reshape(shape)
            o += w
        out["ids"] = np.ascontiguousarray(recs[:, o]).view(np.uint32)
        self.pending = out

    def set_range(self, x0, x1):
        self.range = (x0, x1)

    def histogram(self, hist):
        bx = np.clip(base_x(self.st["x"], self.res), 0, self.res)
        hist += torch.from_numpy(np.bincount(bx, minlength=self.n1).astype(np.int32))

    def num_particles(self):
        return len(self.st["x"]) + (0 if self.pending is None else len(self.pending["x"]))

    def download_slots(self):
        st = self.st
        if self.pending is not None:
            st = {k: np.concatenate([self.st[k], self.pending[k]]) for k in self.st}
        return {k: st[k] for k in ("x", "v", "F", "C", "Jp", "ids")}

    def synchronize(self):
        pass

    def launch_count(self):
        return 0
