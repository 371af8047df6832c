"""Host side of the data-parallel training step (BASELINE config 5): flat parameter / gradient /
Adam-state tensors, the NCCL gradient all-reduce, Keras-style fit loop.

Replaces ``Model.fit_generator`` as used by ``FplNetwork.train`` (flypylib/fplnetwork.py:112-122) and
the tower replication of ``make_train_parallel`` (fplnetwork.py:124-128, flypylib/multi_gpu.py:20-61):
one process per GPU, each rank takes its ``batch_size`` slice of the generator's batch
(``tf.slice`` in multi_gpu.py:21-25), normalises with its own batch statistics, the gradients are
summed over ranks (all-reduce), every rank applies the identical Adam step.
"""
import ctypes

import numpy as np

from . import _lib

ADAM = dict(lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)     # Keras 'adam' defaults
BN_MOMENTUM = 0.99


PRECISIONS = {"fp32": 0, "bf16": 1, "tf32": 2}     # FPL_PREC_* of include/fpl_b200.h


class Trainer(object):
    def __init__(self, model, patch_sz, batch_size, device=None, precision="tf32"):
        """precision: arithmetic of the convolution contractions -- 'tf32' (default: tcgen05 tensor cores on bf16
        hi/lo split operands, fp32-class results), 'bf16' (one bf16 contraction) or 'fp32' (CUDA-core validation
        path)."""
        import torch
        self.model = model
        self.ctx = _lib.context(device)
        self.dev = torch.device("cuda", self.ctx.device)
        self.batch = int(batch_size)
        self.patch = int(patch_sz)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().fpl_train_create(self.ctx.handle, model.spec["id"], self.patch, self.batch,
                                               ctypes.byref(h)), "fpl_train_create")
        self.handle = h
        self.precision = precision
        _lib.check(_lib.lib().fpl_train_set_precision(h, PRECISIONS[precision]), "fpl_train_set_precision")
        n_p, n_bn = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.lib().fpl_train_sizes(h, ctypes.byref(n_p), ctypes.byref(n_bn)))
        flat = np.concatenate([w.ravel() for w in model.get_weights()]).astype(np.float32)
        assert flat.size == n_p.value, (flat.size, n_p.value)
        self.params = torch.from_numpy(flat).to(self.dev)
        self.broadcast_params()
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.bn_batch = torch.zeros(n_bn.value, dtype=torch.float32, device=self.dev)
        self.step_count = 0

    def broadcast_params(self):
        """Every replica must start from rank 0's weights (each rank drew its own glorot initialisation): the
        reference's towers share ONE set of variables (flypylib/multi_gpu.py:20-61)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.params, 0)

    def load_from_model(self):
        """Re-read the flat parameters from the model (after Model.set_weights / a checkpoint restore) and reset the
        Adam moments."""
        import torch
        flat = np.concatenate([w.ravel() for w in self.model.get_weights()]).astype(np.float32)
        self.params.copy_(torch.from_numpy(flat))
        self.broadcast_params()
        self.m.zero_(); self.v.zero_()
        self.step_count = 0

    def forward_backward(self, x, labels, global_batch, seed):
        """x: (B,s,s,s) float32 CUDA tensor, labels: (B,) uint8 CUDA tensor. -> (loss_sum, n_correct)"""
        import torch
        loss, ok = ctypes.c_double(), ctypes.c_int64()
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().fpl_train_forward_backward(
                self.handle, x.data_ptr(), labels.data_ptr(), self.params.data_ptr(), self.grads.data_ptr(),
                self.bn_batch.data_ptr(), 1.0 / float(global_batch), ctypes.c_uint64(seed & ((1 << 64) - 1)),
                ctypes.byref(loss), ctypes.byref(ok), _lib.current_stream_ptr(self.dev.index)),
                "fpl_train_forward_backward")
        return loss.value, ok.value

    def allreduce(self):
        """Sum the tower gradients; average the batch statistics that feed the moving averages."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)
            dist.all_reduce(self.bn_batch, op=dist.ReduceOp.SUM)
            self.bn_batch /= dist.get_world_size()

    def apply(self):
        import torch
        self.step_count += 1
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().fpl_train_apply(
                self.handle, self.params.data_ptr(), self.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                self.bn_batch.data_ptr(), self.step_count, ADAM["lr"], ADAM["beta_1"], ADAM["beta_2"],
                ADAM["epsilon"], BN_MOMENTUM, _lib.current_stream_ptr(self.dev.index)), "fpl_train_apply")

    def train_on_batch(self, data, labels, seed=None):
        """One step on the generator's batch (data (B_total,s,s,s,1), labels (B_total,1,1,1,1)); with
        torch.distributed initialised every rank consumes its own slice.  Returns (loss, accuracy)."""
        import torch
        import torch.distributed as dist
        world, rank = 1, 0
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()
        data = np.asarray(data)
        labels = np.asarray(labels).reshape(data.shape[0])
        global_batch = self.batch * world
        if data.shape[0] != global_batch:
            raise ValueError("generator batch %d != n_gpu*batch_size = %d" % (data.shape[0], global_batch))
        sl = slice(rank * self.batch, (rank + 1) * self.batch)
        x = torch.from_numpy(np.ascontiguousarray(data[sl, ..., 0], dtype=np.float32)).to(self.dev)
        y = torch.from_numpy(np.ascontiguousarray(labels[sl], dtype=np.uint8)).to(self.dev)
        if seed is None:
            seed = 0x5EED0000 + self.step_count * 1000003 + rank
        loss_sum, ok = self.forward_backward(x, y, global_batch, seed)
        self.allreduce()
        self.apply()
        stats = torch.tensor([loss_sum, float(ok)], dtype=torch.float64, device=self.dev)
        if world > 1:
            dist.all_reduce(stats)
        return float(stats[0]) / global_batch, float(stats[1]) / global_batch

    def sync_to_model(self):
        """Write the flat device parameters back into the model (Keras get_weights() layout)."""
        flat = self.params.cpu().numpy()
        out, o = [], 0
        for s in self.model.weight_shapes():
            n = int(np.prod(s))
            out.append(flat[o:o + n].reshape(s).copy())
            o += n
        self.model._set_weights_from_trainer(out)

    def close(self):
        if self.handle:
            _lib.lib().fpl_train_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
