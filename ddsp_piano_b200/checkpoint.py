"""Reader for TensorFlow checkpoints (tensor-bundle format) without TensorFlow.

The reference ships its weights as ``tf.train.Checkpoint`` bundles
(``ddsp_piano/model_weights/{dafx22,v2}/ckpt-*.index`` + ``.data-00000-of-00001``) and restores
them through ``ddsp.training.trainers.Trainer.restore`` (``synthesize_midi_file.py:68``).  The
synthesis path needs two kinds of tensors from them -- the reverb impulse responses
(``reverb_model/reverb_dict/.../embeddings``) and the feedback-delay-network parameters
(``reverb_model/_input_gain/embeddings`` ...) -- so this module parses the bundle directly:

* ``*.index`` is a LevelDB-style sorted string table: data blocks of prefix-compressed
  (key, value) entries, an index block of block handles, a 48-byte footer;
* every value is a ``BundleEntryProto`` {1: dtype, 2: shape, 3: shard_id, 4: offset, 5: size};
* ``*.data-*`` holds the raw little-endian tensors at those offsets.

Host-side utility (NumPy only); nothing here touches the GPU.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8,
           9: np.int64, 10: np.bool_, 17: np.uint16, 22: np.uint32, 23: np.uint64}


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7f) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block(buf, offset, size):
    """Decode one table block into [(key, value)]."""
    data = buf[offset:offset + size]
    if buf[offset + size] != 0:
        raise ValueError('compressed checkpoint index blocks are not supported')
    n_restarts = struct.unpack_from('<I', data, len(data) - 4)[0]
    end = len(data) - 4 - 4 * n_restarts
    pos, key, out = 0, b'', []
    while pos < end:
        shared, pos = _varint(data, pos)
        non_shared, pos = _varint(data, pos)
        vlen, pos = _varint(data, pos)
        key = key[:shared] + bytes(data[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(data[pos:pos + vlen])))
        pos += vlen
    return out


def _fields(msg):
    """Minimal protobuf wire decoder: {field number: [values]} (varints and length-delimited)."""
    pos, out = 0, {}
    while pos < len(msg):
        tag, pos = _varint(msg, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _varint(msg, pos)
        elif wire == 2:
            n, pos = _varint(msg, pos)
            val = msg[pos:pos + n]
            pos += n
        elif wire == 5:
            val = struct.unpack_from('<I', msg, pos)[0]
            pos += 4
        elif wire == 1:
            val = struct.unpack_from('<Q', msg, pos)[0]
            pos += 8
        else:
            raise ValueError(f'unsupported protobuf wire type {wire}')
        out.setdefault(field, []).append(val)
    return out


def _shape(msg):
    dims = []
    for dim in _fields(msg).get(2, []):                 # TensorShapeProto.dim
        size = _fields(dim).get(1, [0])[0]
        dims.append(size - (1 << 64) if size >= (1 << 63) else size)
    return dims


def latest_checkpoint(path):
    """A checkpoint DIRECTORY (the reference's ``--ckpt`` default, ``model_weights/v2/``) resolves to the
    prefix its ``checkpoint`` state file names (what ``tf.train.latest_checkpoint`` returns); a prefix
    is returned unchanged."""
    import os
    import re
    if not os.path.isdir(path):
        return path
    state = os.path.join(path, 'checkpoint')
    if not os.path.exists(state):
        raise FileNotFoundError(f'{path} is a directory without a "checkpoint" state file')
    m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), flags=re.M)
    if not m:
        raise ValueError(f'{state} names no model_checkpoint_path')
    name = m.group(1)
    return name if os.path.isabs(name) else os.path.join(path, name)


class Checkpoint:
    """``Checkpoint(prefix)`` with ``prefix`` like ``.../model_weights/dafx22/ckpt-0``."""

    def __init__(self, prefix):
        prefix = latest_checkpoint(prefix)
        self.prefix = prefix
        with open(prefix + '.index', 'rb') as f:
            buf = f.read()
        if struct.unpack_from('<Q', buf, len(buf) - 8)[0] != _MAGIC:
            raise ValueError(f'{prefix}.index is not a tensor-bundle index')
        footer = buf[len(buf) - 48:]
        _, pos = _varint(footer, 0)                      # metaindex handle (unused)
        _, pos = _varint(footer, pos)
        index_off, pos = _varint(footer, pos)
        index_size, pos = _varint(footer, pos)
        self.entries = {}
        self.n_shards = 1
        for _, handle in _block(buf, index_off, index_size):
            off, p = _varint(handle, 0)
            size, _ = _varint(handle, p)
            for key, value in _block(buf, off, size):
                if key == b'':                           # BundleHeaderProto {1: num_shards}
                    self.n_shards = _fields(value).get(1, [1])[0]
                    continue
                f = _fields(value)
                self.entries[key.decode()] = dict(
                    dtype=f.get(1, [0])[0], shape=_shape(f[2][0]) if 2 in f else [],
                    shard=f.get(3, [0])[0], offset=f.get(4, [0])[0], size=f.get(5, [0])[0])

    def keys(self):
        return sorted(self.entries)

    def find(self, fragment):
        """Keys containing ``fragment`` (variable keys end in '/.ATTRIBUTES/VARIABLE_VALUE')."""
        return [k for k in self.keys() if fragment in k]

    def tensor(self, key):
        if key not in self.entries:
            hits = [k for k in self.find(key) if k.endswith('/.ATTRIBUTES/VARIABLE_VALUE')
                    and '.OPTIMIZER_SLOT' not in k]
            if len(hits) != 1:
                raise KeyError(f'{key!r} matches {len(hits)} variables: {hits[:5]}')
            key = hits[0]
        e = self.entries[key]
        if e['dtype'] not in _DTYPES:
            raise ValueError(f'{key}: unsupported dtype enum {e["dtype"]}')
        dt = np.dtype(_DTYPES[e['dtype']]).newbyteorder('<')
        path = f'{self.prefix}.data-{e["shard"]:05d}-of-{self.n_shards:05d}'
        with open(path, 'rb') as f:
            f.seek(e['offset'])
            raw = f.read(e['size'])
        return np.frombuffer(raw, dtype=dt).reshape(e['shape']).astype(dt.newbyteorder('='))

    # -- what the synthesis path needs ---------------------------------------------------------
    def reverb_ir(self, piano_model=0):
        """Row of the impulse-response embedding (MultiInstrumentReverb, sub_modules.py:351-365)."""
        return self.tensor('reverb_model/reverb_dict')[piano_model]

    def fdn_parameters(self, piano_model=0):
        """Controls of MultiInstrumentFeedbackDelayReverb.call (sub_modules.py:425-441) for one
        instrument, activations applied: relu on the reverberation time, sigmoid on alpha_tone,
        allpass embeddings [32] -> [8, 4] (split in 4, stacked on the last axis)."""
        def emb(name):
            return self.tensor(f'reverb_model/{name}/embeddings')[piano_model].astype(np.float32)
        split = lambda x: np.stack(np.split(x, 4, axis=-1), axis=-1)
        return dict(input_gain=emb('_input_gain'), output_gain=emb('_output_gain'),
                    gain_allpass=split(emb('_gain_allpass')), delays_allpass=split(emb('_delays_allpass')),
                    time_rev_0_sec=np.maximum(emb('_time_rev_0_sec'), 0.0),
                    alpha_tone=(1.0 / (1.0 + np.exp(-emb('_alpha_tone')))).astype(np.float32),
                    early_ir=emb('_early_ir'))


class NpzWeights:
    """The same ``tensor(key)`` lookup over an ``.npz`` export of a checkpoint (keys = the
    checkpoint's variable keys, see ``tests/golden/make_model_weights.py``): lets the model be
    restored where the TensorFlow bundle itself is not available."""

    def __init__(self, path):
        self.path = path
        with np.load(path) as z:
            self.arrays = {k: z[k] for k in z.files}

    def keys(self):
        return sorted(self.arrays)

    def find(self, fragment):
        return [k for k in self.keys() if fragment in k]

    def tensor(self, key):
        if key not in self.arrays:
            hits = [k for k in self.find(key) if k.endswith('/.ATTRIBUTES/VARIABLE_VALUE')]
            if len(hits) != 1:
                raise KeyError(f'{key!r} matches {len(hits)} variables: {hits[:5]}')
            key = hits[0]
        return self.arrays[key]
