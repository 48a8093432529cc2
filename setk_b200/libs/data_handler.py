"""
setk_b200.libs.data_handler -- the readers / writers the adaptive-beamformer CLI
needs, with the behaviour of the reference's scripts/sptk/libs/data_handler.py:

  parse_scps 139-170, Reader 173-236, ScpReader 239-253, WaveReader 326-413,
  NumpyReader 439-448, SpectrogramReader 483-503, ScriptReader 506-535,
  Writer 274-310, WaveWriter 590-605, NumpyWriter 608-622.

Host-side IO is not accelerated (SURVEY.md section 8a "boundary only"); what
changes is SpectrogramReader, whose STFT runs in libsetk_b200 (all channels of
an utterance in one setk_stft launch instead of one librosa call per channel).
Kaldi matrices: uncompressed float / double matrices and vectors (tokens FM, DM,
FV, DV of kaldi_io.py:136-362) are read; compressed matrices (CM*) are not.
"""
import codecs
import glob
import os
import random
import struct
import subprocess
import sys
import warnings
from io import BytesIO, TextIOWrapper
from pathlib import Path

import numpy as np

from .utils import filekey, get_plan, read_wav, write_wav

__all__ = [
    "WaveWriter", "NumpyWriter", "SpectrogramReader", "ScriptReader", "WaveReader",
    "NumpyReader", "ScpReader", "parse_scps"
]


def run_command(command, wait=True):
    """Run a shell pipeline, return (stdout, stderr)  (data_handler.py:30-49)."""
    p = subprocess.Popen(command, shell=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if wait:
        stdout, stderr = p.communicate()
        if p.returncode != 0:
            raise Exception("There was an error while running the command \"{0}\":\n{1}\n".format(
                command, bytes.decode(stderr)))
        return stdout, stderr
    return p


def _fopen(fname, mode):
    """open() extended with "-" (stdin/stdout) and "cmd |" (pipe)  (data_handler.py:73-104)."""
    if mode not in ["w", "r", "wb", "rb"]:
        raise ValueError(f"Unknown open mode: {mode}")
    if not fname:
        return None
    fname = fname.strip()
    if fname == "-":
        if mode in ["w", "wb"]:
            return sys.stdout.buffer if mode == "wb" else sys.stdout
        return sys.stdin.buffer if mode == "rb" else sys.stdin
    if fname[-1] == "|":
        if mode not in ["rb", "r"]:
            raise RuntimeError("Now only support input from pipe")
        p = subprocess.Popen(fname[:-1], shell=True, stdout=subprocess.PIPE)
        return p.stdout if mode == "rb" else TextIOWrapper(p.stdout)
    if mode in ["r", "rb"] and not os.path.exists(fname):
        raise FileNotFoundError(f"Could not find common file: \"{fname}\"")
    if mode in ["r", "w"]:
        return codecs.open(fname, mode, encoding="utf-8")
    return open(fname, mode)


def _fclose(fname, fd):
    if fname != "-" and fd and fname[-1] != "|":
        fd.close()


class ext_open(object):

    def __init__(self, fname, mode):
        self.fname = fname
        self.mode = mode

    def __enter__(self):
        self.fd = _fopen(self.fname, self.mode)
        return self.fd

    def __exit__(self, *args):
        _fclose(self.fname, self.fd)


def parse_scps(scp_path, value_processor=lambda x: x, num_tokens=2, restrict=True):
    """Parse a Kaldi script (.scp) file, stdin / pipe allowed  (data_handler.py:139-170)."""
    line = 0
    scp_dict = {}
    with ext_open(scp_path, "r") as f:
        for raw_line in f:
            scp_tokens = raw_line.strip().split()
            line += 1
            if not scp_tokens:
                continue
            if scp_tokens[-1] == "|":
                key, value = scp_tokens[0], " ".join(scp_tokens[1:])
            else:
                if (num_tokens >= 2 and len(scp_tokens) != num_tokens) or (restrict and
                                                                           len(scp_tokens) < 2):
                    raise RuntimeError(f"For {scp_path}, format error in " +
                                       f"line[{line:d}]: {raw_line}")
                if num_tokens == 2:
                    key, value = scp_tokens
                else:
                    key, value = scp_tokens[0], scp_tokens[1:]
            if key in scp_dict:
                raise ValueError(f"Duplicated key \'{key}\' exists in {scp_path}")
            scp_dict[key] = value_processor(value)
    return scp_dict


class Reader(object):
    """Reader template (data_handler.py:173-236)."""

    def __init__(self, index_dict):
        self.index_dict = index_dict
        self.index_keys = list(self.index_dict.keys())

    def _load(self, key):
        return self.index_dict[key]

    def sample(self, num_items):
        keys = random.sample(self.index_keys, num_items)
        samp = [(key, self._load(key)) for key in keys]
        return samp[0] if num_items == 1 else samp

    def __len__(self):
        return len(self.index_dict)

    def __contains__(self, key):
        return key in self.index_dict

    def __iter__(self):
        for key in self.index_keys:
            yield key, self._load(key)

    def __getitem__(self, index):
        if type(index) not in [int, str]:
            raise IndexError(f"Unsupported index type: {type(index)}")
        if type(index) == int:
            num_utts = len(self.index_keys)
            if index >= num_utts or index < 0:
                raise KeyError("Interger index out of range, " + f"{index:d} vs {num_utts:d}")
            index = self.index_keys[index]
        if index not in self.index_dict:
            raise KeyError(f"Missing utterance {index}!")
        return self._load(index)

    def get(self, index, default=None):
        return self.__getitem__(index) if index in self else default


class ScpReader(Reader):
    """Kaldi's scp reader (data_handler.py:239-253)."""

    def __init__(self, scp_rspecifier, value_processor=lambda x: x, num_tokens=2, restrict=True):
        super(ScpReader, self).__init__(
            parse_scps(scp_rspecifier, value_processor=value_processor, num_tokens=num_tokens,
                       restrict=restrict))


class WaveReader(ScpReader):
    """
    Sequential/Random reader for single/multiple channel wave (data_handler.py:326-413):
    a path, a glob of per-channel files (sorted), `ark:offset`, or a `cmd |` pipe.
    """

    def __init__(self, wav_scp, sr=16000, normalize=True):
        super(WaveReader, self).__init__(wav_scp)
        self.sr = sr
        self.normalize = normalize
        self.wav_ark_mgr = {}

    def read_internal(self, addr, beg=None, end=None):
        if isinstance(addr, str) and ":" in addr:
            tokens = addr.split(":")
            if len(tokens) != 2:
                raise RuntimeError(f"Value format error: {addr}")
            fname, offset = tokens[0], int(tokens[1])
            if fname not in self.wav_ark_mgr:
                self.wav_ark_mgr[fname] = open(fname, "rb")
            wav_ark = self.wav_ark_mgr[fname]
            wav_ark.seek(offset)
            return read_wav(wav_ark, beg=beg or 0, end=end, normalize=self.normalize, sr=self.sr)
        return read_wav(addr, beg=beg or 0, end=end, normalize=self.normalize, sr=self.sr)

    def read(self, key, beg=None, end=None):
        fname = self.index_dict[key].rstrip()
        if fname[-1] == "|":
            stdout, _ = run_command(fname[:-1], wait=True)
            return self.read_internal(BytesIO(stdout))
        wav_list = glob.glob(fname)
        n = len(wav_list)
        if n == 0:
            raise RuntimeError("Could not find file matches " + f"template \'{fname}\'")
        if n == 1:
            return self.read_internal(wav_list[0], beg=beg, end=end)
        # in sorted order, sentitive to beamforming
        return np.vstack([self.read_internal(addr, beg=beg, end=end) for addr in sorted(wav_list)])

    def _load(self, key):
        return self.read(key)

    def maxabs(self, key):
        return np.max(np.abs(self.read(key)))

    def duration(self, key):
        return self.read(key).shape[-1] / self.sr

    def nsamps(self, key):
        return self.read(key).shape[-1]

    def power(self, key):
        samps = self.read(key)
        s = samps if samps.ndim == 1 else samps[0]
        return np.linalg.norm(s, 2)**2 / s.size


class NumpyReader(ScpReader):
    """Reader for numpy's ndarray (*.npy) files (data_handler.py:439-448)."""

    def __init__(self, npy_scp):
        super(NumpyReader, self).__init__(npy_scp)

    def _load(self, key):
        return np.load(self.index_dict[key])


class SpectrogramReader(WaveReader):
    """
    Sequential/Random reader for single/multiple channel STFT (data_handler.py:483-503).
    All channels go through one setk_stft launch; returns numpy complex64
    F x T (mono) or N x F x T, or their transposes with transpose=True.
    """

    def __init__(self, wav_scp, normalize=True, **kwargs):
        super(SpectrogramReader, self).__init__(wav_scp, normalize=normalize)
        self.stft_kwargs = kwargs

    def stft(self, samps, as_tensor=False):
        import torch
        from .utils import default_device
        kw = dict(frame_len=1024, frame_hop=256, round_power_of_two=True, center=False,
                  window="hann", transpose=True)
        kw.update(self.stft_kwargs)
        for unsupported in ("apply_abs", "apply_log", "apply_pow"):
            if kw.pop(unsupported, False):
                raise NotImplementedError(f"SpectrogramReader: {unsupported} is not supported here")
        mono = samps.ndim == 1
        x = np.ascontiguousarray(samps if not mono else samps[None])
        dev = default_device()
        pl = get_plan(x.shape[0], kw["frame_len"], kw["frame_hop"], kw["center"],
                      kw["round_power_of_two"], kw["window"], x.shape[1], dev)
        S = pl.stft(torch.from_numpy(x).to(dev)[None])[0]          # N x F x T
        if kw["transpose"]:
            S = S.transpose(-1, -2)
        if mono:
            S = S[0]
        return S if as_tensor else S.cpu().numpy()

    def _load(self, key):
        return self.stft(super().read(key))


def _read_token(fd):
    tok = b""
    while True:
        c = fd.read(1)
        if c in (b" ", b""):
            break
        tok += c
    return tok.decode()


def read_kaldi_matrix(fd):
    """
    One uncompressed Kaldi float/double matrix or vector from a binary stream
    positioned at the "\\0B" marker (the value an scp's `ark:offset` points at).
    Tokens as in the reference's kaldi_io.py:136-362.
    """
    if fd.read(2) != b"\0B":
        raise RuntimeError("Kaldi object is not in binary mode")
    tok = _read_token(fd)
    if tok[:2] == "CM":
        raise NotImplementedError("compressed Kaldi matrices are not supported")
    if tok not in ("FM", "DM", "FV", "DV"):
        raise RuntimeError(f"Unknown Kaldi object token: {tok}")
    dtype = np.float32 if tok[0] == "F" else np.float64

    def read_int():
        size = struct.unpack("b", fd.read(1))[0]
        if size != 4:
            raise RuntimeError("Kaldi int32 size marker expected")
        return struct.unpack("<i", fd.read(4))[0]

    if tok[1] == "M":
        rows, cols = read_int(), read_int()
        data = np.frombuffer(fd.read(rows * cols * np.dtype(dtype).itemsize), dtype=dtype)
        return data.reshape(rows, cols).copy()
    n = read_int()
    return np.frombuffer(fd.read(n * np.dtype(dtype).itemsize), dtype=dtype).copy()


class ScriptReader(ScpReader):
    """Reader for Kaldi scripts of BaseFloat matrices (data_handler.py:506-535)."""

    def __init__(self, ark_scp):

        def addr_processor(addr):
            addr_token = addr.split(":")
            if len(addr_token) == 1:
                raise ValueError("Unsupported scripts address format")
            return (":".join(addr_token[0:-1]), int(addr_token[-1]))

        super(ScriptReader, self).__init__(ark_scp, value_processor=addr_processor)
        self.fmgr = dict()

    def _load(self, key):
        path, addr = self.index_dict[key]
        if path not in self.fmgr:
            self.fmgr[path] = open(path, "rb")
        fd = self.fmgr[path]
        fd.seek(addr)
        return read_kaldi_matrix(fd)


class Writer(object):
    """Basic writer (data_handler.py:274-310)."""

    def __init__(self, obj_path_or_dir, scp_path=None, is_dir=False):
        self.scp_path = scp_path
        if obj_path_or_dir == "-" and scp_path:
            warnings.warn("Ignore script output discriptor cause dump archives to stdout")
            self.scp_path = None
        self.dump_out_dir = is_dir
        if is_dir:
            self.path_or_dir = Path(obj_path_or_dir).absolute()
            self.path_or_dir.mkdir(exist_ok=True, parents=True)
        else:
            self.path_or_dir = os.path.abspath(obj_path_or_dir)

    def __enter__(self):
        if not self.dump_out_dir:
            self.ark_file = _fopen(self.path_or_dir, "wb")
        self.scp_file = _fopen(self.scp_path, "w")
        return self

    def __exit__(self, *args):
        if not self.dump_out_dir:
            _fclose(self.path_or_dir, self.ark_file)
        _fclose(self.scp_path, self.scp_file)

    def check_args(self, data):
        if not isinstance(data, np.ndarray):
            raise RuntimeError("Instance of Writer accepts np.ndarray object, " +
                               f"but got {type(data)}")

    def write(self, key, data):
        raise NotImplementedError


class WaveWriter(Writer):
    """Writer for wave files (data_handler.py:590-605)."""

    def __init__(self, dump_dir, scp_path=None, sr=16000, normalize=True):
        super(WaveWriter, self).__init__(dump_dir, scp_path, is_dir=True)
        self.sr = sr
        self.normalize = normalize

    def write(self, key, obj):
        self.check_args(obj)
        obj_path = self.path_or_dir / f"{key}.wav"
        write_wav(obj_path, obj, sr=self.sr, normalize=self.normalize)
        if self.scp_file:
            self.scp_file.write(f"{key}\t{obj_path}\n")


class NumpyWriter(Writer):
    """Writer for numpy ndarray (data_handler.py:608-622)."""

    def __init__(self, dump_dir, scp_path=None):
        super(NumpyWriter, self).__init__(dump_dir, scp_path, is_dir=True)

    def write(self, key, obj):
        self.check_args(obj)
        obj_path = self.path_or_dir / f"{key}.npy"
        np.save(obj_path, obj)
        if self.scp_file:
            self.scp_file.write(f"{key}\t{obj_path}\n")
