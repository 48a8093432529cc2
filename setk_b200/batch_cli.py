"""
setk_b200.batch_cli -- the batched feeder behind scripts/sptk/apply_adaptive_beamformer.py.

The reference scales by forking `nj` processes over wav.scp shards, each running one utterance
at a time (scripts/run_adapt_beamformer.sh:66-92, data_handler.py:326-393).  One B200 wants
hundreds of utterances per launch, so the drop-in CLI's offline path is organised as

    reader thread   wav.scp / mask tables -> (key, PCM-16 or float samples C x N, mask T x F)
    main thread     sorts a look-ahead window by (channels, length), cuts it into batches of
                    <= batch_size utterances, fills PINNED staging buffers (ragged: n_samples)
                    and submits them to a 2-lane HostBatchStreamer: H2D, int16/32768 on the
                    device, fused STFT+cov -> fp64 weights -> fused apply+iSTFT+peak-norm+
                    floor(y*32768), D2H of int16 -- copies of batch i+1 overlap kernels of i
    writer thread   waits for a lane's completion event, maps solver status bits to the
                    reference's per-utterance LinAlgError skip (apply_adaptive_beamformer.py:
                    170-172) and writes <dst_dir>/<key>.wav (PCM-16) through WaveWriter

Under torchrun (WORLD_SIZE > 1) every rank takes the keys  rank::world  of the scp on its own
GPU -- split_scp.pl's round robin -- with no communication (results are files).
"""
import os
import queue
import threading
import time

import numpy as np
import torch

from .engine import BeamformPipeline, HostBatchStreamer, status_message
from . import _lib


class _Staging(object):
    """One set of pinned host buffers for a batch (two sets alternate)."""

    def __init__(self, batch, C, n_cap, t_cap, F, n_out_cap, pcm16, with_itf):
        self.audio = torch.empty((batch, C, n_cap), dtype=torch.int16 if pcm16 else torch.float32,
                                 pin_memory=True)
        self.mask = torch.zeros((batch, t_cap, F), dtype=torch.float32, pin_memory=True)
        self.mask_n = torch.zeros((batch, t_cap, F), dtype=torch.float32, pin_memory=True) \
            if with_itf else None
        self.n_samples = torch.zeros((batch,), dtype=torch.int32, pin_memory=True)
        self.out = torch.empty((batch, n_out_cap), dtype=torch.int16, pin_memory=True)
        self.status = torch.zeros((batch,), dtype=torch.int32, pin_memory=True)


class BatchedBeamformer(object):
    """
    Offline mask-based beamforming of a stream of utterances in ragged batches.
    make_pipeline(C, n_cap, batch) -> BeamformPipeline.
    """

    def __init__(self, make_pipeline, num_bins, plan_probe, batch_size=64, lookahead=256, device=None,
                 logger=None):
        self.make_pipeline = make_pipeline
        self.num_bins = num_bins
        self.plan_probe = plan_probe           # (C, n) -> (frames, output samples)
        self.batch_size = int(batch_size)
        self.lookahead = max(int(lookahead), self.batch_size)
        self.device = device
        self.log = logger
        self._streamer = None
        self._key = None                       # (C, n_cap, pcm16, with_itf)
        self._sets = []
        self._free = queue.Queue()
        self.stats = {"batches": 0, "utterances": 0, "padded_samples": 0, "samples": 0,
                      "t_fill": 0.0, "t_wait": 0.0}

    # -- capacity management -------------------------------------------------
    def _ensure(self, C, n_max, pcm16, with_itf):
        if self._streamer is not None and self._key[0] == C and self._key[1] >= n_max and \
                self._key[2] == pcm16 and self._key[3] == with_itf:
            return
        self.drain()
        n_cap = int(n_max) if self._streamer is None else int(n_max * 1.25)
        n_cap = (n_cap + 3) & ~3
        self._streamer = None
        self._sets = []
        torch.cuda.empty_cache()
        self._streamer = HostBatchStreamer(lambda: self.make_pipeline(C, n_cap, self.batch_size),
                                           self.batch_size, C, n_cap, slots=2, pcm16=pcm16,
                                           pcm16_out=True, device=self.device,
                                           run_kwargs={"normalize": True})
        t_cap, n_out_cap = self.plan_probe(C, n_cap)
        self._key = (C, n_cap, pcm16, with_itf)
        self._free = queue.Queue()
        for _ in range(2):
            st = _Staging(self.batch_size, C, n_cap, t_cap, self.num_bins, n_out_cap, pcm16, with_itf)
            self._sets.append(st)
            self._free.put(st)

    def drain(self):
        """Wait until every submitted batch has been written."""
        if self._streamer is None:
            return
        got = [self._free.get() for _ in self._sets]
        for st in got:
            self._free.put(st)

    # -- one batch -------------------------------------------------------------
    def submit(self, items, results):
        """
        items: list of (key, samps (C,N) int16|float32, mask (T,F) f32, itf mask or None).
        results: queue the writer thread reads (lane, staging, keys, out lengths).
        """
        C = items[0][1].shape[0]
        pcm16 = items[0][1].dtype == np.int16
        with_itf = items[0][3] is not None
        n_max = max(it[1].shape[1] for it in items)
        self._ensure(C, n_max, pcm16, with_itf)
        t0 = time.perf_counter()
        st = self._free.get()                    # blocks until the writer released a set
        self.stats["t_wait"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        keys, lens = [], []
        a_np, m_np = st.audio.numpy(), st.mask.numpy()
        mn_np = st.mask_n.numpy() if with_itf else None
        for b, (key, samps, mask, itf) in enumerate(items):
            n = samps.shape[1]
            T_b, n_out = self.plan_probe(C, n)
            if mask.shape != (T_b, self.num_bins):
                raise ValueError(f"{key}: mask shape {mask.shape} does not match the STFT "
                                 f"({T_b} x {self.num_bins})")
            a_np[b, :, :n] = samps
            m_np[b, :T_b] = mask
            if with_itf:
                mn_np[b, :T_b] = itf
            st.n_samples[b] = n
            keys.append(key)
            lens.append(n_out)
            self.stats["samples"] += n
            self.stats["padded_samples"] += self._key[1]
        self.stats["t_fill"] += time.perf_counter() - t0
        lane = self._streamer.submit(st.audio, st.mask, st.out, n_samples=st.n_samples, count=len(items),
                                     h_mask_n=st.mask_n, h_status=st.status)
        ev = torch.cuda.Event()
        ev.record(lane["stream"])
        results.put((ev, st, keys, lens))
        self.stats["batches"] += 1
        self.stats["utterances"] += len(items)

    def release(self, st):
        self._free.put(st)


def plan_batches(window, batch_size):
    """Sort a look-ahead window by (channels, dtype, has itf, length) and cut it into batches."""
    window.sort(key=lambda it: (it[1].shape[0], str(it[1].dtype), it[3] is not None, it[1].shape[1]))
    batches, cur = [], []
    for it in window:
        if cur and (len(cur) >= batch_size or it[1].shape[0] != cur[0][1].shape[0] or
                    it[1].dtype != cur[0][1].dtype or (it[3] is None) != (cur[0][3] is None)):
            batches.append(cur)
            cur = []
        cur.append(it)
    if cur:
        batches.append(cur)
    return batches


def run_batched(args, wave_reader, tgt_mask_reader, itf_mask_reader, stft_kwargs, num_bins, dev, logger):
    """The offline (non-VAD, non-online) path of the CLI.  Returns the number of utterances written."""
    from .libs.data_handler import WaveWriter
    from .libs.utils import get_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    keys = [k for k in wave_reader.index_keys]
    if world > 1:
        keys = keys[rank::world]

    def make_pipeline(C, n_cap, batch):
        return BeamformPipeline(C, beamformer=args.beamformer,
                                round_power_of_two=bool(args.round_power_of_two), ban=bool(args.ban),
                                pmwf_ref=args.pmwf_ref, rank1_appro=args.rank1_appro,
                                post_masking=bool(args.mask), max_batch=batch, max_samples=n_cap,
                                device=dev, **stft_kwargs)

    probes = {}

    def plan_probe(C, n):
        pl = probes.get(C)
        if pl is None:
            pl = probes[C] = get_plan(C, args.frame_len, args.frame_hop, bool(args.center),
                                      bool(args.round_power_of_two), args.window, n, dev)
        T_b = pl.num_frames(n)
        return T_b, pl.istft_length(T_b)

    feeder = BatchedBeamformer(make_pipeline, num_bins, plan_probe, batch_size=args.batch_size,
                               lookahead=args.lookahead, device=dev, logger=logger)
    items_q = queue.Queue(maxsize=feeder.lookahead * 2)
    results = queue.Queue()
    errors = []

    # Loading an utterance is file I/O + numpy (both release the GIL): a small pool of loader threads
    # keeps the order of the scp and hides the per-file latency (one reader thread fed 250 utt/s from
    # a RAM disk; the device wants two orders of magnitude more).  Readers that seek in a shared
    # archive handle (`archive:offset` values, Kaldi scripts) are serialised by a lock.
    import collections
    import contextlib
    from concurrent.futures import ThreadPoolExecutor
    archive_lock = threading.Lock()

    def guarded(rd, key):
        spec = rd.index_dict.get(key)
        shared = not isinstance(spec, str) or ":" in spec or spec.rstrip().endswith("|")
        return archive_lock if shared else contextlib.nullcontext()

    def load(key):
        if key not in tgt_mask_reader:
            return None
        with guarded(wave_reader, key):
            samps = wave_reader[key]
        if samps.ndim == 1:
            samps = samps[None]
        with guarded(tgt_mask_reader, key):
            mask = np.asarray(tgt_mask_reader[key], dtype=np.float32)
        itf = None
        if itf_mask_reader is not None:
            with guarded(itf_mask_reader, key):
                itf = np.asarray(itf_mask_reader[key], dtype=np.float32)
        # make sure the masks are T x F (apply_adaptive_beamformer.py:150-158)
        if mask.shape[0] == num_bins and mask.shape[1] != num_bins:
            mask = np.ascontiguousarray(mask.T)
            if itf is not None:
                itf = np.ascontiguousarray(itf.T)
        first = samps[0].astype(np.float32)
        if samps.dtype == np.int16:
            first = first / np.float32(32768.0)
        power = float(np.linalg.norm(first, 2)**2 / first.size)
        return (key, np.ascontiguousarray(samps), mask, itf), power

    def reader():
        n_loaders = max(1, int(getattr(args, "reader_threads", 8)))
        try:
            with ThreadPoolExecutor(max_workers=n_loaders) as pool:
                pending = collections.deque()

                def emit(fut):
                    got = fut.result()
                    if got is not None:
                        item, power = got
                        logger.info(f"Processing utterance {item[0]}, " +
                                    f"signal power {10 * np.log10(power + 1e-5):.2f}...")
                        items_q.put(item)

                for key in keys:
                    pending.append(pool.submit(load, key))
                    if len(pending) >= 4 * n_loaders:
                        emit(pending.popleft())
                while pending:
                    emit(pending.popleft())
        except BaseException as e:       # surfaced by the main thread
            errors.append(e)
        finally:
            items_q.put(None)

    num_done = [0]

    def writer_loop(writer):
        while True:
            job = results.get()
            if job is None:
                return
            ev, st, bkeys, lens = job
            try:
                ev.synchronize()
                status = st.status.numpy()
                out = st.out.numpy()
                for b, key in enumerate(bkeys):
                    bits = int(status[b]) & _lib.ST_ERROR_MASK
                    if bits:
                        # the reference's per-utterance `except np.linalg.LinAlgError` (lines 170-172)
                        logger.error(f"Raise linalg error: {key} ({status_message(bits)})")
                        continue
                    writer.write(key, out[b, :lens[b]].copy())
                    num_done[0] += 1
            except BaseException as e:
                errors.append(e)
            finally:
                feeder.release(st)

    with WaveWriter(args.dst_dir, sr=args.sr) as writer:
        rt = threading.Thread(target=reader, daemon=True)
        wt = threading.Thread(target=writer_loop, args=(writer,), daemon=True)
        rt.start()
        wt.start()
        window, eof = [], False
        t0 = time.perf_counter()
        try:
            while not eof or window:
                while not eof and len(window) < feeder.lookahead:
                    it = items_q.get()
                    if it is None:
                        eof = True
                    else:
                        window.append(it)
                if errors:
                    raise errors[0]
                batches = plan_batches(window, feeder.batch_size)
                # keep a partial last batch for the next window unless the input is exhausted
                window = []
                if not eof and batches and len(batches[-1]) < feeder.batch_size:
                    window = batches.pop()
                for batch in batches:
                    feeder.submit(batch, results)
                    if errors:
                        raise errors[0]
            feeder.drain()
        finally:
            results.put(None)
            wt.join()
        if errors:
            raise errors[0]
        dt = time.perf_counter() - t0
    s = feeder.stats
    if s["utterances"]:
        logger.info(f"Batched feeder: {s['utterances']} utterances in {s['batches']} batches, "
                    f"{dt:.2f} s ({s['utterances'] / max(dt, 1e-9):.1f} utts/s), padding "
                    f"{100.0 * (1 - s['samples'] / max(1, s['padded_samples'])):.1f} %, "
                    f"fill {s['t_fill']:.2f} s, waiting for the device/writer {s['t_wait']:.2f} s")
    return num_done[0]
