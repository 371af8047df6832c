"""CPU: the HDF5 reader / writer behind ``load_network`` / ``Model.save`` / h5 path inputs (flypylib_b200/h5lite.py;
reference call sites flypylib/fplnetwork.py:32-44,81-97,137-139 and fplobjdetect.py:154-156 -- h5py / Keras are absent).

The reader is pinned against the one genuine HDF5 file on the image (written by MATLAB through libhdf5: user block,
superblock 0, symbol-table group, float64 dataset with a string attribute) and round-trips everything the writer emits."""
import os
import pickle

import numpy as np
import pytest

from flypylib_b200 import h5lite

MAT = None
try:
    import scipy.io
    MAT = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
except Exception:                                   # pragma: no cover
    pass


@pytest.mark.skipif(MAT is None or not os.path.exists(MAT), reason="scipy's HDF5 test file is not installed")
def test_reads_a_file_written_by_libhdf5():
    f = h5lite.File(MAT)
    assert f.base == 512 and f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.float64
    assert np.allclose(d[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
    assert d.attrs["MATLAB_class"] == b"double"


def test_round_trip_tree(tmp_path):
    rng = np.random.default_rng(3)
    tree = {"attrs": {"title": "a tree", "count": np.int64(7), "vec": np.arange(5, dtype=np.float32)},
            "items": {"g%02d" % i: {"attrs": {"i": np.int32(i)},
                                   "items": {"x": rng.standard_normal((3, i + 1)).astype(np.float32),
                                             "deep": {"items": {"y": {"data": np.arange(i, dtype=np.int16),
                                                                      "attrs": {"unit": b"vox"}}}}}}
                      for i in range(20)}}
    tree["items"]["u8"] = rng.integers(0, 255, (4, 5, 6)).astype(np.uint8)
    tree["items"]["f64"] = rng.standard_normal((2, 2))
    p = str(tmp_path / "t.h5")
    h5lite.write_h5(p, tree)
    f = h5lite.File(p)
    assert f.attrs["title"] == b"a tree" and f.attrs["count"] == 7 and np.array_equal(f.attrs["vec"], np.arange(5))
    assert sorted(f.keys()) == sorted(tree["items"])
    for i in range(20):
        g = f["g%02d" % i]
        assert g.attrs["i"] == i
        assert np.array_equal(g["x"][:], tree["items"]["g%02d" % i]["items"]["x"])
        y = f["/g%02d/deep/y" % i]
        assert np.array_equal(y[:], np.arange(i, dtype=np.int16)) and y.attrs["unit"] == b"vox"
    assert np.array_equal(f["u8"][:], tree["items"]["u8"]) and f["u8"].dtype == np.uint8
    assert np.array_equal(f["f64"][:], tree["items"]["f64"])
    with pytest.raises(KeyError):
        f["missing"]


@pytest.mark.parametrize("gzip,shuffle", [(None, False), (4, False), (6, True)])
def test_chunked_volume(tmp_path, gzip, shuffle):
    """``/main`` volumes as h5py writes them with chunks / compression (edge chunks are padded)."""
    vol = np.random.default_rng(5).integers(0, 255, (21, 17, 30)).astype(np.uint8)
    f32 = np.random.default_rng(6).random((9, 10, 11)).astype(np.float32)
    p = str(tmp_path / "v.h5")
    h5lite.write_h5(p, {"items": {"main": {"data": vol, "chunks": (8, 8, 16), "gzip": gzip, "shuffle": shuffle},
                                  "pred": {"data": f32, "chunks": (4, 10, 6), "gzip": gzip, "shuffle": shuffle}}})
    f = h5lite.File(p)
    assert np.array_equal(f["/main"][:], vol) and np.array_equal(f["pred"][:], f32)


def test_keras_weight_files_round_trip(tmp_path):
    """Model.save -> Keras layout (/model_weights/<layer>/<layer>/<weight>:0, layer_names / weight_names) -> load_weights."""
    from flypylib_b200 import fplmodels
    from oracle import models_oracle as M
    for arch in ("vgg_like", "vgg_like2", "unet_like2"):
        model = getattr(fplmodels, arch)()[0]
        w = M.random_weights(arch, seed=5)
        model.set_weights(w)
        p = str(tmp_path / (arch + ".h5"))
        model.save(p)
        f = h5lite.File(p)
        names = [n.decode() for n in f["model_weights"].attrs["layer_names"]]
        assert names[0] == "conv3d_1" and names[1] == "batch_normalization_1"
        assert f["model_weights/conv3d_1/conv3d_1/kernel:0"].shape == w[0].shape
        arrays, wnames = h5lite.read_keras_weights(p)
        assert len(arrays) == len(w) and all(np.array_equal(a, b) for a, b in zip(arrays, w))
        assert wnames[1] == "batch_normalization_1/batch_normalization_1/gamma:0"
        other = getattr(fplmodels, arch)()[0]
        other.load_weights(p)
        assert all(np.array_equal(a, b) for a, b in zip(other.get_weights(), w))
        model.save_weights(p)                                  # save_weights layout: the groups sit at the root
        other = getattr(fplmodels, arch)()[0]
        other.load_weights(p)
        assert all(np.array_equal(a, b) for a, b in zip(other.get_weights(), w))
    wrong = fplmodels.vgg_like()[0]
    with pytest.raises(ValueError):
        wrong.load_weights(p)                                  # unet_like2 weights into a vgg_like


def test_load_network_reads_a_reference_style_pickle(tmp_path):
    """A pickle that names ``flypylib.fplnetwork.FplNetwork`` / ``flypylib.fplmodels.vgg_like`` (what the reference's
    save_network writes, fplnetwork.py:81-90) + ``<path>.keras.h5`` load through ``load_network`` without GPU work."""
    from flypylib_b200 import fplmodels, fplnetwork
    from oracle import models_oracle as M
    net = fplnetwork.FplNetwork(fplmodels.vgg_like)
    w = M.random_weights("vgg_like", seed=11)
    net.train_single.set_weights(w)
    path = str(tmp_path / "net.p")
    net.save_network(path)
    assert os.path.exists(path + ".keras.h5")
    blob = pickle.dumps(net, protocol=0).replace(b"flypylib_b200", b"flypylib")
    state = net.__getstate__()
    assert b"cflypylib.fplnetwork\nFplNetwork" in blob and state["train_single"] is None
    with open(path, "wb") as fh:
        fh.write(blob)
    back = fplnetwork.load_network(path)
    assert type(back) is fplnetwork.FplNetwork and back.model is fplmodels.vgg_like
    assert back.rf_size == net.rf_size and back.infer_sz == net.infer_sz
    assert all(np.array_equal(a, b) for a, b in zip(back.train_single.get_weights(), w))
    assert all(np.array_equal(a, b) for a, b in zip(back.infer_network.get_weights(), w))


def test_truncated_and_foreign_files(tmp_path):
    p = str(tmp_path / "x.h5")
    with open(p, "wb") as fh:
        fh.write(b"not an hdf5 file" * 10)
    with pytest.raises(h5lite.H5Error):
        h5lite.File(p)
    h5lite.write_h5(p, {"items": {"main": np.arange(1000, dtype=np.float32)}})
    data = open(p, "rb").read()
    with open(p, "wb") as fh:
        fh.write(data[:len(data) // 3])
    with pytest.raises((h5lite.H5Error, ValueError, IndexError, KeyError)):
        h5lite.File(p)["main"][:]
