"""Drop-in for the model builders of flypylib/fplmodels.py on the T-bar inference path.

``vgg_like`` (:102-136), ``vgg_like2`` (:138-172) and ``unet_like2`` (:258-304) keep the reference
contract: ``builder(in_sz=None) -> (model, (rf_size, rf_offset, rf_stride), infer_sz, compile_args)``.
``model`` is not a Keras graph but a thin handle on a B200 network object (C ABI ``fpl_net_*``); it
offers the part of the Keras ``Model`` API that ``FplNetwork`` uses: ``summary``, ``compile``,
``get_weights`` / ``set_weights`` (Keras ``get_weights()`` order and layouts), ``input_shape``,
``predict(x, batch_size)``.
"""
import ctypes

import numpy as np

from . import _lib
from . import fplutils

# (kind, ...) rows purely for summary()/weight bookkeeping; the executable graph lives in csrc/net.cu
_ARCH = {
    "vgg_like": dict(id=_lib.ARCH_VGG_LIKE, rf=(18, 7, 4), infer_sz=102, final_bias=True,
                     convs=[(3, 1, 48), (1, 48, 48), (3, 48, 48), (1, 48, 48), (3, 48, 48), (1, 48, 96),
                            (1, 96, 96)], final_cin=96),
    "vgg_like2": dict(id=_lib.ARCH_VGG_LIKE2, rf=(24, 10, 4), infer_sz=100, final_bias=True,
                      convs=[(3, 1, 48), (3, 48, 48), (3, 48, 48), (3, 48, 48), (3, 48, 48), (1, 48, 96),
                             (1, 96, 96)], final_cin=96),
    "unet_like2": dict(id=_lib.ARCH_UNET_LIKE2, rf=(24, 9, 1), infer_sz=100, final_bias=False,
                       convs=[(3, 1, 32), (3, 32, 32), (3, 32, 64), (3, 64, 64), (1, 64, 128), (3, 192, 64),
                              (1, 64, 64), (3, 96, 32), (1, 32, 32)], final_cin=32),
    "baseline_model": dict(id=_lib.ARCH_BASELINE, rf=(18, 7, 4), infer_sz=102, final_bias=True,
                           convs=[(3, 1, 32), (3, 32, 32), (3, 32, 32), (1, 32, 64)], final_cin=64),
    # fplmodels.py:174-208; a 4th entry False marks a convolution without BatchNormalization (the shortcut)
    "resnet_like": dict(id=_lib.ARCH_RESNET_LIKE, rf=(18, 7, 4), infer_sz=102, final_bias=True,
                        convs=[(3, 1, 32), (3, 32, 32), (1, 32, 32), (3, 32, 64), (1, 32, 64, False), (1, 64, 64)],
                        final_cin=64),
    # fplmodels.py:470-526: no BatchNormalization (all convolutions marked False)
    "unet_like_vol": dict(id=_lib.ARCH_UNET_LIKE_VOL, rf=(62, 6, 1), infer_sz=102, final_bias=False,
                          convs=[(3, 1, 16, False), (1, 16, 16, False), (3, 16, 32, False), (1, 32, 32, False),
                                 (1, 32, 64, False), (3, 96, 64, False), (1, 64, 64, False), (3, 80, 32, False),
                                 (1, 32, 32, False)], final_cin=32),
    "unet_like": dict(id=_lib.ARCH_UNET_LIKE, rf=(18, 6, 1), infer_sz=102, final_bias=False,
                      convs=[(3, 1, 32), (1, 32, 32), (3, 32, 64), (1, 64, 64), (1, 64, 128), (3, 192, 64),
                             (1, 64, 64), (3, 96, 32), (1, 32, 32)], final_cin=32),
    "unet_like3": dict(id=_lib.ARCH_UNET_LIKE3, rf=(32, 13, 1), infer_sz=100, final_bias=False,
                       convs=[(3, 1, 32), (3, 32, 32), (3, 32, 64), (3, 64, 64), (3, 64, 128), (1, 128, 128),
                              (3, 192, 64), (1, 64, 64), (3, 96, 32), (1, 32, 32)], final_cin=32),
    "unet_like4": dict(id=_lib.ARCH_UNET_LIKE4, rf=(40, 17, 1), infer_sz=100, final_bias=False,
                       convs=[(3, 1, 32), (3, 32, 32), (3, 32, 64), (3, 64, 64), (3, 64, 128), (3, 128, 128),
                              (3, 192, 64), (1, 64, 64), (3, 96, 32), (1, 32, 32)], final_cin=32),
    "unet_like4b": dict(id=_lib.ARCH_UNET_LIKE4B, rf=(40, 17, 1), infer_sz=100, final_bias=False,
                        convs=[(3, 1, 32), (3, 32, 32), (3, 32, 64), (1, 64, 32), (3, 32, 64), (1, 64, 48),
                               (3, 48, 128), (1, 128, 48), (3, 48, 128), (1, 128, 48), (3, 112, 64), (1, 64, 64),
                               (3, 96, 32), (1, 32, 32)], final_cin=32),
}

_PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "tf32": _lib.PREC_TF32}
DEFAULT_PRECISION = "bf16"


def masked_focal_loss(y_true, y_pred):      # named so compile_args keep the reference's keys (:45-50)
    raise NotImplementedError("training losses are outside the B200 inference hot path")


def masked_accuracy(y_true, y_pred):
    raise NotImplementedError("training metrics are outside the B200 inference hot path")


def lb0l1err(y_true, y_pred):
    raise NotImplementedError("training metrics are outside the B200 inference hot path")


def lb1l1err(y_true, y_pred):
    raise NotImplementedError("training metrics are outside the B200 inference hot path")


def _rank():
    try:
        import torch.distributed as dist
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    except ImportError:
        return 0


class Model(object):
    """Handle on one network: weights on the host in Keras order + a lazily created device net."""

    def __init__(self, arch, in_sz, upsample_output=False, precision=None):
        self.arch = arch
        self.spec = _ARCH[arch]
        in_sz = fplutils.to3d(in_sz) if in_sz is not None else (None, None, None)
        self.input_shape = (None,) + tuple(in_sz) + (1,)
        self.upsample_output = upsample_output
        self.precision = precision or DEFAULT_PRECISION
        self.compile_args = None
        self._weights = self._initial_weights()
        self._net = None
        self._net_dirty = True
        self._device = None

    # ---- Keras-like surface ------------------------------------------------------------------
    def weight_shapes(self):
        shapes = []
        for cv in self.spec["convs"]:
            k, cin, cout = cv[:3]
            shapes.append((k, k, k, cin, cout))
            if len(cv) < 4 or cv[3]:
                shapes += [(cout,)] * 4
        shapes.append((1, 1, 1, self.spec["final_cin"], 1))
        if self.spec["final_bias"]:
            shapes.append((1,))
        return shapes

    def _initial_weights(self):
        """Keras defaults: glorot_uniform kernels, BN gamma=1 beta=0 mean=0 var=1, zero bias."""
        ws = []
        for cv in self.spec["convs"]:
            k, cin, cout = cv[:3]
            lim = np.sqrt(6.0 / (k ** 3 * cin + k ** 3 * cout))
            ws.append(np.random.uniform(-lim, lim, (k, k, k, cin, cout)).astype(np.float32))
            if len(cv) < 4 or cv[3]:
                ws += [np.ones(cout, np.float32), np.zeros(cout, np.float32), np.zeros(cout, np.float32),
                       np.ones(cout, np.float32)]
        cin = self.spec["final_cin"]
        lim = np.sqrt(6.0 / (cin + 1))
        ws.append(np.random.uniform(-lim, lim, (1, 1, 1, cin, 1)).astype(np.float32))
        if self.spec["final_bias"]:
            ws.append(np.zeros(1, np.float32))
        return ws

    def get_weights(self):
        return [w.copy() for w in self._weights]

    def set_weights(self, weights):
        shapes = self.weight_shapes()
        if len(weights) != len(shapes):
            raise ValueError("expected %d weight arrays, got %d" % (len(shapes), len(weights)))
        new = []
        for w, s in zip(weights, shapes):
            w = np.ascontiguousarray(w, dtype=np.float32)
            if w.shape != s:
                raise ValueError("weight shape %s does not match %s" % (w.shape, s))
            new.append(w)
        self._weights = new
        self._net_dirty = True
        if getattr(self, "_trainer", None) is not None:      # weights set from outside: the trainer's flat copy
            self._trainer.load_from_model()                  # (and its Adam moments) would be stale

    def _set_weights_from_trainer(self, weights):
        """Trainer -> model write-back at the end of an epoch: same as set_weights, but the trainer stays as it is."""
        tr, self._trainer = getattr(self, "_trainer", None), None
        try:
            self.set_weights(weights)
        finally:
            self._trainer = tr

    def count_params(self):
        return int(sum(int(np.prod(s)) for s in self.weight_shapes()))

    def compile(self, **kwargs):
        self.compile_args = kwargs

    def summary(self):
        print("Model %s  input %s  precision %s" % (self.arch, self.input_shape, self.precision))
        for i, cv in enumerate(self.spec["convs"]):
            k, cin, cout = cv[:3]
            bn = len(cv) < 4 or cv[3]
            print("  conv%d  Conv3D(%d,(%d,%d,%d))%s   in=%d  params=%d"
                  % (i, cout, k, k, k, " + BN" if bn else "", cin, k ** 3 * cin * cout + (4 * cout if bn else 0)))
        print("  predictions Conv3D(1,(1,1,1)) sigmoid  in=%d" % self.spec["final_cin"])
        if self.upsample_output and self.spec["rf"][2] != 1:
            print("  UpSampling3D(%d)" % self.spec["rf"][2])
        print("Total params: %d" % self.count_params())

    # ---- device side -------------------------------------------------------------------------
    def set_precision(self, precision):
        if precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
        if precision != self.precision:
            self.precision = precision
            self._net_dirty = True

    def device_net(self, device=None):
        """fpl_net handle with the current weights uploaded (created on first use)."""
        import torch
        ctx = _lib.context(device)
        lib = _lib.lib()
        if self._net is None or self._device != ctx.device:
            self.close()
            h = ctypes.c_void_p()
            _lib.check(lib.fpl_net_create(ctx.handle, self.spec["id"], ctypes.byref(h)), "fpl_net_create")
            self._net, self._device, self._net_dirty = h, ctx.device, True
        if self._net_dirty:
            arr = (ctypes.c_void_p * len(self._weights))(*[w.ctypes.data for w in self._weights])
            with torch.cuda.device(ctx.device):
                _lib.check(lib.fpl_net_set_weights(self._net, arr, len(self._weights),
                                                   _PRECISIONS[self.precision]), "fpl_net_set_weights")
            self._net_dirty = False
        return self._net

    def out_size(self, in_sz):
        o = ctypes.c_int32()
        _lib.check(_lib.lib().fpl_net_out_size(self.device_net(), int(in_sz), ctypes.byref(o)), "fpl_net_out_size")
        return o.value

    def predict_device(self, tiles):
        """tiles: CUDA float32 tensor (N,s,s,s) -> CUDA float32 (N,o,o,o) (x rf_stride up-sampled)."""
        import torch
        n, s = int(tiles.shape[0]), int(tiles.shape[1])
        dev = tiles.device.index
        net = self.device_net(dev)
        o = self.out_size(s)
        out = torch.empty((n, o, o, o), dtype=torch.float32, device=tiles.device)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fpl_net_forward_tiles(net, tiles.data_ptr(), n, s, out.data_ptr(),
                                                        _lib.current_stream_ptr(dev)), "fpl_net_forward_tiles")
        return out

    def predict(self, x, batch_size=32, verbose=0):
        """Keras Model.predict: x (N,s,s,s,1) -> (N,o,o,o,1) float32 (fplnetwork.py:175-176)."""
        import torch
        x = np.asarray(x)
        if x.ndim != 5 or x.shape[-1] != 1 or not (x.shape[1] == x.shape[2] == x.shape[3]):
            raise ValueError("expected input of shape (N,s,s,s,1), got %s" % (x.shape,))
        _lib.context()
        outs = []
        bs = max(1, int(batch_size))
        for i in range(0, x.shape[0], bs):
            xb = torch.from_numpy(np.ascontiguousarray(x[i:i + bs, ..., 0], dtype=np.float32)).cuda()
            ob = self.predict_device(xb)
            if not self.upsample_output and self.spec["rf"][2] != 1:
                s = self.spec["rf"][2]
                ob = ob[:, ::s, ::s, ::s]
            outs.append(ob.cpu().numpy()[..., None])
        return np.concatenate(outs, 0)

    # ---- training (VGG builders, default binary_crossentropy + adam) ---------------------------
    def configure_training(self, n_gpu, batch_size, input_shape):
        """make_train_parallel bookkeeping: ranks, per-rank batch, patch edge."""
        self._train_cfg = (int(n_gpu), int(batch_size), int(fplutils.to3d(input_shape)[0]))
        self._trainer = None

    def trainer(self):
        from . import fpltrain
        cfg = getattr(self, "_train_cfg", None)
        if cfg is None:
            raise RuntimeError("call FplNetwork.make_train_parallel(n_gpu, batch_size, input_shape) first")
        if getattr(self, "_trainer", None) is None:
            self._trainer = fpltrain.Trainer(self, cfg[2], cfg[1])
        return self._trainer

    def fit_generator(self, generator, steps_per_epoch, epochs, callbacks=None, verbose=1):
        """Keras Model.fit_generator as FplNetwork.train uses it (fplnetwork.py:120-121): per epoch
        steps_per_epoch batches from `generator`; callbacks receive on_epoch_end(epoch, logs)."""
        if self.arch not in ("vgg_like", "vgg_like2"):
            raise NotImplementedError("the B200 training step covers the VGG builders")
        tr = self.trainer()
        history = []
        for epoch in range(int(epochs)):
            loss_sum = acc_sum = 0.0
            for _ in range(int(steps_per_epoch)):
                data, labels = next(generator)
                loss, acc = tr.train_on_batch(data, labels)
                loss_sum += loss; acc_sum += acc
            logs = {"acc": acc_sum / steps_per_epoch, "loss": loss_sum / steps_per_epoch}
            history.append(logs)
            tr.sync_to_model()
            if verbose:
                print("Epoch %d/%d - loss: %.4f - acc: %.4f" % (epoch + 1, epochs, logs["loss"], logs["acc"]))
            if _rank() == 0:                # one log / checkpoint writer (the replicas hold identical weights)
                for cb in (callbacks or []):
                    cb.on_epoch_end(epoch, logs)
        return history

    def keras_layers(self):
        """[(layer name, [(weight name, array), ...])] in Keras layer order: what Model.save writes under
        /model_weights (conv3d_i: kernel [, bias]; batch_normalization_i: gamma, beta, moving_mean, moving_variance)."""
        out, ws, ci, bi = [], list(self._weights), 0, 0
        i = 0
        while i < len(ws):
            ci += 1
            name = "conv3d_%d" % ci
            entry = [("%s/kernel:0" % name, ws[i])]
            i += 1
            if i < len(ws) and ws[i].ndim == 1 and ws[i].shape[0] == entry[0][1].shape[4] and \
                    (i + 1 == len(ws) or ws[i + 1].ndim == 5):
                entry.append(("%s/bias:0" % name, ws[i]))          # the final convolution's bias
                i += 1
                out.append((name, entry))
                continue
            out.append((name, entry))
            if i + 3 < len(ws) and all(w.ndim == 1 for w in ws[i:i + 4]):
                bi += 1
                bn = "batch_normalization_%d" % bi
                out.append((bn, [("%s/%s:0" % (bn, n), ws[i + j])
                                 for j, n in enumerate(("gamma", "beta", "moving_mean", "moving_variance"))]))
                i += 4
        return out

    def save(self, path):
        """Keras ``Model.save`` (flypylib/fplnetwork.py:17,83): an HDF5 file with the weights in the Keras layout
        (``/model_weights/<layer>/<layer>/<weight>:0``, attributes ``layer_names`` / ``weight_names``) written by
        flypylib_b200.h5lite; ``*.npz`` paths keep the array-list container of round 1."""
        if str(path).endswith(".npz"):
            np.savez(path, *self._weights)
            return
        from . import h5lite
        h5lite.write_keras_weights(str(path), self.keras_layers(), full_model=True,
                                   model_config='{"class_name": "Model", "config": {"name": "%s"}}' % self.arch)

    def save_weights(self, path):
        from . import h5lite
        h5lite.write_keras_weights(str(path), self.keras_layers(), full_model=False)

    def load_weights(self, path):
        """Weights from a Keras ``.h5`` file (``Model.save`` or ``save_weights`` layout; layer order of the file =
        ``get_weights()`` order) or from an ``.npz`` array list."""
        if str(path).endswith(".npz"):
            with np.load(path) as z:
                self.set_weights([z['arr_%d' % i] for i in range(len(z.files))])
            return
        from . import h5lite
        arrays, names = h5lite.read_keras_weights(str(path))
        shapes = self.weight_shapes()
        if len(arrays) != len(shapes) or any(tuple(a.shape) != tuple(s) for a, s in zip(arrays, shapes)):
            raise ValueError("%s does not hold the weights of %s: got %s" % (path, self.arch,
                                                                             [tuple(a.shape) for a in arrays]))
        self.set_weights(arrays)

    def close(self):
        if getattr(self, "_trainer", None) is not None:
            self._trainer.close()
            self._trainer = None
        if self._net is not None:
            _lib.lib().fpl_net_destroy(self._net)
            self._net = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def vgg_like(in_sz=None):
    """returns standard model based on VGG architecture (flypylib/fplmodels.py:102-136)"""
    return Model("vgg_like", in_sz), (18, 7, 4), 102, None


def vgg_like2(in_sz=None):
    """returns standard model based on VGG architecture (flypylib/fplmodels.py:138-172)"""
    return Model("vgg_like2", in_sz), (24, 10, 4), 100, None


def unet_like2(in_sz=24):
    """construct a u-net style network (flypylib/fplmodels.py:258-304)"""
    compile_args = {'loss': masked_focal_loss,
                    'optimizer': 'adam',
                    'metrics': [masked_accuracy, lb0l1err, lb1l1err]}
    return Model("unet_like2", in_sz), (24, 9, 1), 100, compile_args


def masked_binary_crossentropy(y_true, y_pred):     # fplmodels.py:52-60, named for compile_args only
    raise NotImplementedError("training losses are outside the B200 inference hot path")


def baseline_model(in_sz=None):
    """returns simple baseline model (flypylib/fplmodels.py:73-100)"""
    return Model("baseline_model", in_sz), (18, 7, 4), 102, None


def masked_weighted_binary_crossentropy(y_true, y_pred):     # fplmodels.py, named for compile_args only
    raise NotImplementedError("training losses are outside the B200 inference hot path")


def unet_like_vol(in_sz=62):
    """fplmodels.py:470-526: U-Net without BatchNormalization (16/32/64 channels), volume-to-volume training."""
    compile_args = {'loss': masked_weighted_binary_crossentropy, 'optimizer': 'adam', 'metrics': ['masked_accuracy']}
    return Model("unet_like_vol", in_sz), (62, 6, 1), 102, compile_args


def resnet_like(in_sz=None):
    """fplmodels.py:174-208: residual blocks (add of a cropped shortcut, BN before the add, ReLU after it)."""
    return Model("resnet_like", in_sz), (18, 7, 4), 102, None


def unet_like(in_sz=18):
    """construct a u-net style network (flypylib/fplmodels.py:206-256)"""
    compile_args = {'loss': masked_binary_crossentropy, 'optimizer': 'adam',
                    'metrics': [masked_accuracy, lb0l1err, lb1l1err]}
    return Model("unet_like", in_sz), (18, 6, 1), 102, compile_args


def unet_like3(in_sz=32):
    """construct a u-net style network (flypylib/fplmodels.py:306-357)"""
    compile_args = {'loss': masked_focal_loss, 'optimizer': 'adam',
                    'metrics': [masked_accuracy, lb0l1err, lb1l1err]}
    return Model("unet_like3", in_sz), (32, 13, 1), 100, compile_args


def unet_like4(in_sz=40):
    """construct a u-net style network (flypylib/fplmodels.py:359-410)"""
    compile_args = {'loss': masked_focal_loss, 'optimizer': 'adam',
                    'metrics': [masked_accuracy, lb0l1err, lb1l1err]}
    return Model("unet_like4", in_sz), (40, 17, 1), 100, compile_args


def unet_like4b(in_sz=40):
    """construct a u-net style network (flypylib/fplmodels.py:412-467)"""
    compile_args = {'loss': masked_focal_loss, 'optimizer': 'adam',
                    'metrics': [masked_accuracy, lb0l1err, lb1l1err]}
    return Model("unet_like4b", in_sz), (40, 17, 1), 100, compile_args
