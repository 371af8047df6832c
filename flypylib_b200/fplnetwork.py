"""Drop-in for flypylib/fplnetwork.py on the inference path: ``FplNetwork`` with the same
constructor, attributes and ``infer`` / ``make_infer_parallel`` methods (reference :46-189).

``infer`` keeps the reference semantics exactly -- tile origins ``k*(infer_sz-2*rf_offset)``,
zero-padded far-edge tiles, scatter of tile interiors, ``rf_offset``-wide border left at 0 -- but the
tiles are gathered, evaluated and scattered on the GPU by ``fpl_net_infer_volume`` (no float64
staging batch, no host loops).
"""
import ctypes
import pickle

import numpy as np

from . import _lib
from . import fplutils
from . import fplmodels


class CSVLogger(object):
    """Keras CSVLogger stand-in: epoch,acc,loss rows (fplnetwork.py:115)."""

    def __init__(self, filename):
        self.filename = filename
        self._started = False

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        keys = sorted(logs)
        with open(self.filename, "a" if self._started else "w") as f:
            if not self._started:
                f.write(",".join(["epoch"] + keys) + "\n")
            f.write(",".join([str(epoch)] + ["%r" % logs[k] for k in keys]) + "\n")
        self._started = True


class multi_gpu_callback(object):
    """fplnetwork.py:9-17: save the single (un-replicated) model at the end of every epoch."""

    def __init__(self, model, save_prefix):
        self.model_to_save = model
        self.save_prefix = save_prefix

    def on_epoch_end(self, epoch, logs=None):
        self.model_to_save.save('%s_%03d.h5' % (self.save_prefix, epoch))


def get_custom_objects(compile_args):
    """fplnetwork.py:19-30: name -> function map of the non-string loss and metrics in compile_args (what
    Keras' load_model needs as custom_objects; kept so that scripts calling it keep working)."""
    custom_objects = {}
    if not isinstance(compile_args['loss'], str):
        custom_objects[compile_args['loss'].__name__] = compile_args['loss']
    for mm in compile_args['metrics']:
        if not isinstance(mm, str):
            custom_objects[mm.__name__] = mm
    return custom_objects


def _weights_path(filepath):
    """Weight file next to the pickled network: ``<filepath>.keras.h5`` as in the reference (fplnetwork.py:82-83),
    read and written by flypylib_b200.h5lite (h5py is not needed).  ``<filepath>.keras.npz`` (round-1 container) is
    still read when no .h5 file exists."""
    import os
    h5 = filepath + '.keras.h5'
    if os.path.exists(h5) or not os.path.exists(filepath + '.keras.npz'):
        return h5
    return filepath + '.keras.npz'


class _RefUnpickler(pickle.Unpickler):
    """Network pickles written by the reference name ``flypylib.fplnetwork.FplNetwork`` and the builder / loss
    functions of ``flypylib.fplmodels``: resolve them to this package."""

    def find_class(self, module, name):
        if module == 'flypylib' or module.startswith('flypylib.'):
            module = 'flypylib_b200' + module[len('flypylib'):]
        return super().find_class(module, name)


def load_network(filepath):
    """fplnetwork.py:32-44: un-pickle an FplNetwork written by ``save_network`` (this package's or the reference's)
    and restore its weights from ``<filepath>.keras.h5``."""
    with open(filepath, 'rb') as fn:
        network = _RefUnpickler(fn).load()
    network._restore_models(_weights_path(filepath))
    return network


class FplNetwork:
    """deep learning/CNN class wrapping a B200 network (reference: wraps a keras model)

    supports full stack inference; training (fit_generator) is outside the B200 hot path.
    """

    def __init__(self, model):
        self.model = model

        self.train_network, rf_info, infer_sz, compile_args = self.model()
        self.train_network.summary()
        self.train_single = self.train_network

        self.rf_size = fplutils.to3d(rf_info[0])
        self.rf_offset = fplutils.to3d(rf_info[1])
        self.rf_stride = fplutils.to3d(rf_info[2])

        self.infer_network = None
        self.n_gpu = 1

        self.infer_sz = fplutils.to3d(infer_sz)

        if compile_args is None:
            compile_args = {'loss': 'binary_crossentropy',
                            'optimizer': 'adam',
                            'metrics': ['accuracy']}
        self.train_network.compile(**compile_args)
        self.compile_args = compile_args
        self.tile_multiplier = 1

    # ------------------------------------------------------------------------------------------
    def save_network(self, filepath):
        """fplnetwork.py:81-97: weights to the side file, the network object (builder, receptive-field info,
        compile_args, ...) pickled to ``filepath``; the live networks are kept."""
        self.train_single.save(filepath + '.keras.h5')
        with open(filepath, 'wb') as fn:
            pickle.dump(self, fn)

    def __getstate__(self):
        keep = dict(self.__dict__)
        keep['_precision'] = self.train_single.precision if self.train_single is not None else None
        for k in ('train_network', 'train_single', 'infer_network', '_copy_stream'):
            keep[k] = None
        return keep

    def _restore_models(self, weights_file):
        precision = self.__dict__.pop('_precision', None)
        self.__dict__.setdefault('tile_multiplier', 1)      # absent from pickles written by the reference
        self.train_single, _, _, _ = self.model()
        self.train_network = self.train_single
        if precision:
            self.train_single.set_precision(precision)
        self.train_single.load_weights(weights_file)
        self.train_network.compile(**self.compile_args)
        self._set_infer()

    # ------------------------------------------------------------------------------------------
    def _set_infer(self):
        """fplnetwork.py:99-110: rebuild at infer_sz (+ UpSampling3D(rf_stride) when strided) and
        copy the trained weights."""
        net, _, _, _ = self.model(self.infer_sz)
        net.upsample_output = True
        net.set_precision(self.train_single.precision)
        self.infer_network = net
        self.infer_network.set_weights(self.train_single.get_weights())

    def set_precision(self, precision):
        """'bf16' (tcgen05, default), 'tf32' (tcgen05) or 'fp32' (CUDA-core validation path)."""
        self.train_single.set_precision(precision)
        if self.infer_network is not None:
            self.infer_network.set_precision(precision)

    def train(self, generator, steps_per_epoch, epochs,
              log_file, save_filepath):
        """fplnetwork.py:112-122: CSV log of the epoch metrics, per-epoch save of the single model,
        fit over the generator, then rebuild the inference network with the trained weights."""
        csv_logger = CSVLogger(log_file)
        checkpoint = multi_gpu_callback(self.train_single, save_filepath)
        callbacks = [csv_logger, checkpoint]
        if not hasattr(self.train_network, "_train_cfg"):
            self.make_train_parallel(1, 64, self.rf_size)
        self.train_network.fit_generator(
            generator, steps_per_epoch, epochs, callbacks=callbacks)
        self._set_infer()

    def make_train_parallel(self, n_gpu, batch_size, input_shape):
        """fplnetwork.py:124-128.  n_gpu towers of batch_size patches each = n_gpu ranks (one process per
        GPU, torch.distributed/NCCL) that all-reduce their gradients; see fpltrain.Trainer."""
        self.train_network = self.train_single
        self.train_network.configure_training(n_gpu, batch_size, input_shape)
        self.train_network.compile(**self.compile_args)

    def make_infer_parallel(self, n_gpu):
        """fplnetwork.py:130-134.  The reference replicates the graph on n_gpu towers inside one process
        (flypylib/multi_gpu.py:20-61) and feeds one tile per tower and predict step; here n_gpu is the number
        of ranks (one process per GPU, torch.distributed) over which ``infer`` shards the volume as z-slabs
        (``multi_gpu.shard_plan`` + ``infer_slab_device``).  With n_gpu > 1 ``torch.distributed`` must be
        initialised with that world size before ``infer`` is called."""
        self._set_infer()
        self.n_gpu = n_gpu

    # ------------------------------------------------------------------------------------------
    def _check_built(self):
        assert self.infer_network is not None, 'network has not been trained'
        assert self.infer_network.input_shape[1:-1] == self.infer_sz, \
            'network input shape does not match expected infer_sz'

    def slab_granularity(self):
        """Plane granularity of z cuts that keep ``infer`` bit-identical to the whole-volume call: rf_stride for
        the VGG builders (shift-equivariant), the reference tile pitch for the U-Nets (tile phase matters)."""
        stride = int(self.rf_stride[0])
        return stride if stride != 1 else int(self.infer_sz[0]) - 2 * int(self.rf_offset[0])

    def infer_slab_device(self, image_slab, Z, z0, normalize=None, pred=None, pred_z0=None):
        """Planes [z0, z0 + len(image_slab)) of a (Z,Y,X) volume as one slab (fpl_net_infer_slab): writes the
        prediction planes [z0+off, z1-off) -- plus the zero border planes when the slab touches an end of the
        volume -- into ``pred``, a CUDA float32 tensor whose plane 0 is plane ``pred_z0`` of the prediction volume
        (default: a fresh tensor covering exactly the written planes).  Returns (pred, first, last) with
        [first,last) the written plane range."""
        import torch
        self._check_built()
        if image_slab.dim() != 3:
            raise ValueError("image slab must be 3-D (z,Y,X)")
        is_u8 = image_slab.dtype == torch.uint8
        if not is_u8 and image_slab.dtype != torch.float32:
            image_slab = image_slab.float()
        if is_u8 and normalize is None:
            raise ValueError("uint8 input needs normalize=(mean, std)")
        image_slab = image_slab.contiguous()
        zs, Y, X = (int(v) for v in image_slab.shape)
        Z, z0 = int(Z), int(z0)
        z1 = z0 + zs
        off = int(self.rf_offset[0])
        first = 0 if z0 == 0 else z0 + off
        last = Z if z1 == Z else z1 - off
        if last < first:
            last = first
        if pred is None:
            pred = torch.empty((last - first, Y, X), dtype=torch.float32, device=image_slab.device)
            pred_z0 = first
        if pred_z0 is None:
            pred_z0 = 0
        if not (pred.is_contiguous() and pred.dtype == torch.float32 and tuple(pred.shape[1:]) == (Y, X)):
            raise ValueError("pred must be a contiguous float32 (planes,Y,X) tensor")
        if first < pred_z0 or last > pred_z0 + int(pred.shape[0]):
            raise ValueError("pred does not cover the written planes [%d,%d)" % (first, last))
        dev = image_slab.device.index
        net = self.infer_network.device_net(dev)
        lib = _lib.lib()
        _lib.check(lib.fpl_net_set_tile_multiplier(net, int(self.tile_multiplier)), "fpl_net_set_tile_multiplier")
        mean, std = (float(normalize[0]), float(normalize[1])) if normalize is not None else (0.0, 1.0)
        # address of plane z0 of the prediction volume inside `pred` (only [first,last) is touched)
        p_slab = pred.data_ptr() + (z0 - int(pred_z0)) * Y * X * 4
        with torch.cuda.device(dev):
            _lib.check(lib.fpl_net_infer_slab(net, image_slab.data_ptr(), 1 if is_u8 else 0, mean, std, Z, z0, z1,
                                              Y, X, ctypes.c_void_p(p_slab), _lib.current_stream_ptr(dev)),
                       "fpl_net_infer_slab")
        return pred, first, last

    def infer_device(self, image_dev, normalize=None, out=None):
        """``infer`` on a CUDA tensor (Z,Y,X): float32 (already normalised) or uint8 with
        ``normalize=(mean, std)`` applied on the fly.  Returns a CUDA float32 tensor."""
        import torch
        self._check_built()
        if image_dev.dim() != 3:
            raise ValueError("image must be 3-D (Z,Y,X)")
        Z, Y, X = (int(s) for s in image_dev.shape)
        pred = out if out is not None else torch.empty((Z, Y, X), dtype=torch.float32, device=image_dev.device)
        self.infer_slab_device(image_dev, Z, 0, normalize=normalize, pred=pred, pred_z0=0)
        return pred

    def infer_host(self, image_host, normalize=None, out=None, image_dev=None, chunk_layers=4):
        """``infer`` on a HOST tensor (pinned memory for true overlap) with the host->device copy pipelined
        behind the computation: the volume is cut into chunks of ``chunk_layers`` reference tile layers in z;
        chunk c+1 is copied on a side stream while chunk c runs.  Every chunk is a slab whose first plane lies
        on the reference tile grid (``infer_slab_device``), so it writes exactly the planes of the whole-volume
        ``infer`` straight into the result -- no staging buffer, no extra device copies.  Returns a CUDA float32
        tensor.  ``image_dev`` (optional) receives the device copy of the volume."""
        import torch
        self._check_built()
        if image_host.dim() != 3:
            raise ValueError("image must be 3-D (Z,Y,X)")
        dev = torch.device("cuda", torch.cuda.current_device())
        Z, Y, X = (int(v) for v in image_host.shape)
        off, out_sz = int(self.rf_offset[0]), int(self.infer_sz[0]) - 2 * int(self.rf_offset[0])
        n_layers = 0 if Z <= 2 * off else -(-(Z - 2 * off) // out_sz)
        if image_dev is None:
            image_dev = torch.empty((Z, Y, X), dtype=image_host.dtype, device=dev)
        pred = out if out is not None else torch.empty((Z, Y, X), dtype=torch.float32, device=dev)
        if n_layers <= chunk_layers:
            image_dev.copy_(image_host, non_blocking=True)
            return self.infer_device(image_dev, normalize=normalize, out=pred)
        side = getattr(self, "_copy_stream", None)
        if side is None:
            side = self._copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        side.wait_stream(main)                      # image_dev / pred may still be in use by earlier work
        chunks, copied = [], 0
        for k0 in range(0, n_layers, chunk_layers):
            k1 = min(n_layers, k0 + chunk_layers)
            in0, in1 = k0 * out_sz, (Z if k1 == n_layers else k1 * out_sz + 2 * off)
            with torch.cuda.stream(side):
                if in1 > copied:
                    image_dev[copied:in1].copy_(image_host[copied:in1], non_blocking=True)
                    copied = in1
                ev = side.record_event()
            chunks.append((in0, in1, ev))
        for in0, in1, ev in chunks:
            main.wait_event(ev)
            self.infer_slab_device(image_dev[in0:in1], Z, in0, normalize=normalize, pred=pred, pred_z0=0)
        return pred

    def infer(self, image):
        """fplnetwork.py:136-189: probability map (float32, image.shape) of a 3-D image.

        With ``make_infer_parallel(n_gpu > 1)`` and torch.distributed initialised (one process per GPU) every
        rank evaluates its z-slab and the prediction planes are all-gathered, so every rank returns the whole map
        -- bit-identical to the single-GPU result."""
        import torch
        if isinstance(image, str):                       # fplnetwork.py:137-139: h5 file with the volume in /main
            from . import h5lite
            image = h5lite.File(image)['/main'][:]
        self._check_built()
        world = self._world()
        if isinstance(image, torch.Tensor):
            dev = image if image.is_cuda else image.cuda()
            return self.infer_device(dev) if world == 1 else self._infer_gathered(dev)
        image = np.asarray(image)
        _lib.context()
        if world > 1:
            return self._infer_gathered(image).cpu().numpy()
        dev = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).cuda()
        return self.infer_device(dev).cpu().numpy()

    def _world(self):
        if self.n_gpu <= 1:
            return 1
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("make_infer_parallel(%d): initialise torch.distributed (one process per GPU) "
                               "before calling infer" % self.n_gpu)
        if dist.get_world_size() != self.n_gpu:
            raise RuntimeError("make_infer_parallel(%d) but the process group has %d ranks"
                               % (self.n_gpu, dist.get_world_size()))
        return self.n_gpu

    def _infer_gathered(self, image):
        """Sharded infer + all-gather of the prediction planes (every rank passes the same whole image)."""
        import torch
        import torch.distributed as dist
        from . import multi_gpu
        world, rank = dist.get_world_size(), dist.get_rank()
        Z, Y, X = (int(v) for v in image.shape)
        plans = multi_gpu.shard_plan(Z, int(self.rf_offset[0]), self.slab_granularity(), world)
        (in0, in1), (own0, own1) = plans[rank]
        dev = torch.device("cuda", torch.cuda.current_device())
        if isinstance(image, torch.Tensor):
            slab = image[in0:in1].to(dev)
            if slab.dtype != torch.float32:
                slab = slab.float()
        else:
            slab = torch.from_numpy(np.ascontiguousarray(image[in0:in1], dtype=np.float32)).to(dev)
        own = torch.zeros((0, Y, X), dtype=torch.float32, device=dev)
        if in1 > in0:
            own, first, last = self.infer_slab_device(slab, Z, in0)
            assert (first, last) == (own0, own1), ((first, last), (own0, own1))
        width = max(p[1][1] - p[1][0] for p in plans)
        pad = torch.zeros((max(width, 1), Y, X), dtype=torch.float32, device=dev)
        pad[:own.shape[0]] = own
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        pred = torch.empty((Z, Y, X), dtype=torch.float32, device=dev)
        for part, (_, (o0, o1)) in zip(parts, plans):
            pred[o0:o1] = part[:o1 - o0]
        return pred
