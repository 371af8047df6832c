"""Detection-list wire formats: the json a ``voxel2obj`` result is written to and
read back from on either side of the hot path.  Host-side, plain Python.

Mirrors the format functions of flypylib/fplsynapses.py:11-111 (``load_from_json``,
``tbars_to_json_format``, ``tbars_to_json_format_raveler``).  The DVID push/pull helpers of that
module (:113 onwards) are storage/network control plane and out of scope.
"""
import json
import os

import numpy as np

from . import fplutils


def _conf_str(c):
    return '%.03f' % c


def tbars_to_json_format(tbars_np, json_file=None, user_name='$fpl', labels=None):
    """DVID annotation elements for a ``{'locs','conf'}`` dict (flypylib/fplsynapses.py:77-96):
    one ``{'Kind': 'PreSyn', 'Pos': [x,y,z] ints (truncated), 'Prop': {'conf': '%.03f', 'user'}}``
    per detection, plus ``'body ID'`` when ``labels`` is given.  Written to ``json_file`` if set."""
    locs, conf = tbars_np['locs'], tbars_np['conf']
    elements = []
    for i in range(int(np.size(conf))):
        el = {'Kind': 'PreSyn',
              'Pos': locs[i, :].astype('int').tolist(),
              'Prop': {'conf': _conf_str(conf[i]), 'user': user_name}}
        if labels is not None:
            el['body ID'] = str(labels[i])
        elements.append(el)
    if json_file is not None:
        with open(json_file, 'w') as f:
            json.dump(elements, f)
    return elements


def tbars_to_json_format_raveler(tbars_np, json_file=None):
    """Raveler layout (flypylib/fplsynapses.py:98-111): ``{'data': [{'T-bar': {'confidence': '%.03f',
    'location': [x,y,z]}}, ...]}``."""
    locs, conf = tbars_np['locs'], tbars_np['conf']
    rows = [{'T-bar': {'confidence': _conf_str(conf[i]), 'location': locs[i, :].astype('int').tolist()}}
            for i in range(int(np.size(conf)))]
    doc = {'data': rows}
    if json_file is not None:
        with open(json_file, 'w') as f:
            json.dump(doc, f)
    return doc


def load_from_json(fn, vol_sz=None, buffer=None):
    """Read either format back (flypylib/fplsynapses.py:11-75).  ``fn`` is a path, or the json text
    itself when no such file exists.  Raveler documents keep the confidence as stored (a string when
    written by ``tbars_to_json_format_raveler``); DVID elements keep only ``PreSyn`` kinds, parse
    ``Prop['conf']`` to float (default 1.0) and ``Prop['err']`` (default None).  With ``buffer`` (and
    ``vol_sz``), detections closer than ``buffer`` to a face of the volume are dropped.
    Returns ``{'locs', 'conf', 'err'}`` arrays."""
    if os.path.isfile(fn):
        with open(fn) as f:
            data = json.load(f)
    else:
        data = json.loads(fn)
    locs, conf, err = [], [], []
    if isinstance(data, dict) and 'data' in data:
        for syn in data['data']:
            locs.append(syn['T-bar']['location'])
            conf.append(syn['T-bar']['confidence'])
    elif data is not None:
        if len(data) == 1 and isinstance(data[0], list):
            data = data[0]
        for syn in data:
            if syn['Kind'] != 'PreSyn':
                continue
            prop = syn['Prop']
            conf.append(float(prop['conf']) if 'conf' in prop else 1.0)
            err.append(float(prop['err']) if 'err' in prop else None)
            locs.append(syn['Pos'])
    locs, conf, err = np.asarray(locs), np.asarray(conf), np.asarray(err)
    if locs.size > 0 and buffer is not None and buffer != 0:
        assert vol_sz is not None, 'to apply buffer, must also supply volume size'
        b, sz = fplutils.to3d(buffer), fplutils.to3d(vol_sz)
        drop = np.zeros(locs.shape[0], dtype=bool)
        for ax in range(3):
            drop |= (locs[:, ax] < b[ax]) | (locs[:, ax] >= sz[ax] - b[ax])
        locs, conf = locs[~drop], conf[~drop]
    return {'locs': locs, 'conf': conf, 'err': err}
