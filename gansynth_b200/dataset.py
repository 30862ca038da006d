"""NSynth input pipeline (reference dataset.py:12-91) without tf.data.

    nsynth_input_fn(filenames, batch_size, num_epochs, shuffle, buffer_size=None, pitches=None, sources=None)
        -> (waveforms [B, 64000] float32, one-hot labels [B, len(pitches)] float32)

Same name, arguments and stage order as the reference: TFRecordDataset -> shuffle(buffer, reshuffled every
epoch) -> repeat(num_epochs) -> parse Example {path, pitch, source} + read_file + decode_wav(1 channel,
64000 samples) -> filter(min(pitches) <= pitch <= max(pitches), source in sources) -> batch(drop_remainder)
-> prefetch(1).  The reference calls the function once and gets graph tensors that produce a new batch per
`session.run`; here every call with the same arguments returns the NEXT batch of one cached pipeline (so
`functools.partial(nsynth_input_fn, ...)` is the `real_input_fn` GANSynth expects, gan_synth_main.py:58-68)
and raises StopIteration when the epochs are exhausted (tf.errors.OutOfRangeError).

Data path: a background thread (prefetch) has the native library read and decode the batch's WAV files on
os.cpu_count() host threads straight into a pinned int16 buffer (gs_wav_read_batch), copies it to the device
on a side stream and converts int16 -> float32 / 32768 there (gs_pcm16_to_float): 128 KB per clip cross PCIe
instead of 256 KB, and the host never touches a float sample.  The filter runs on the parsed (pitch, source)
BEFORE the audio is read -- same batches as the reference (its predicate ignores the waveform), none of the
wasted decodes.
"""
import ctypes
import os
import queue
import threading

import numpy as np
import torch

from . import _lib
from . import tfrecord

WAVEFORM_LENGTH = 64000      # dataset.py:35 desired_samples


def read_index(filenames, verify=True):
    """All (path, pitch, source) records of the TFRecord files, in file order (make_tfrecord.py:27-47)."""
    paths, pitch, source = [], [], []
    for fn in filenames:
        for rec in tfrecord.read_records(fn, verify=verify):
            ex = tfrecord.parse_example(rec)
            paths.append(ex["path"][0].decode("utf-8"))
            pitch.append(int(ex["pitch"][0]))
            source.append(int(ex["source"][0]))
    return paths, np.asarray(pitch, np.int64), np.asarray(source, np.int64)


def decode_wav_files(paths, desired_samples=WAVEFORM_LENGTH, out=None, threads=None):
    """tf.read_file + audio_ops.decode_wav(desired_channels=1, desired_samples) for a list of files -> int16
    [len(paths), desired_samples] (torch tensor, written in place when `out` is given)."""
    n = len(paths)
    if out is None:
        out = torch.empty((n, desired_samples), dtype=torch.int16)
    assert out.dtype == torch.int16 and out.is_contiguous() and out.shape[0] >= n and out.shape[1] == desired_samples
    arr = (ctypes.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    status = (ctypes.c_int * max(n, 1))()
    _lib.host_call("gs_wav_read_batch", arr, n, out.data_ptr(), desired_samples, threads or os.cpu_count() or 1, status)
    return out


class NSynthPipeline(object):
    """Iterator over (waveforms, labels) batches; see the module docstring for the stage order.
    device="cpu" (host-logic tests) stops before the device half: it yields the decoded int16 batch, there is
    no host-side float conversion."""

    def __init__(self, filenames, batch_size, num_epochs, shuffle, buffer_size=None, pitches=None, sources=None,
                 device=None, seed=0, threads=None, prefetch=1, waveform_length=WAVEFORM_LENGTH):
        if device is None:
            device = "cuda"
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())   # the producer thread pins this device
        self.batch_size = int(batch_size)
        self.num_epochs = num_epochs
        self.shuffle = bool(shuffle)
        self.buffer_size = buffer_size
        self.pitches = sorted(pitches) if pitches else None       # index_table_from_tensor(sorted(pitches)), :15
        self.sources = list(sources) if sources else None
        self.waveform_length = int(waveform_length)
        self.threads = threads or os.cpu_count() or 1
        self.seed = seed
        self.paths, self.pitch, self.source = read_index(list(filenames))
        self._keep = self._filter_mask()
        self._label_of = {p: i for i, p in enumerate(self.pitches)} if self.pitches else {}
        self._records = self._record_stream()
        self._stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._queue = queue.Queue(maxsize=max(1, int(prefetch)))
        self._error = None
        self._done = False
        self._closed = False
        self._thread = threading.Thread(target=self._producer, daemon=True)
        self._thread.start()

    # -------------------------------------------------------------------------------------------- stages
    def _filter_mask(self):
        """dataset.py:65-75."""
        keep = np.ones(len(self.paths), bool)
        if self.pitches:
            keep &= (self.pitch >= min(self.pitches)) & (self.pitch <= max(self.pitches))
        if self.sources:
            keep &= np.isin(self.source, np.asarray(self.sources))
        return keep

    def _epoch_order(self, epoch):
        n = len(self.paths)
        if not self.shuffle:
            return np.arange(n)
        rng = np.random.default_rng([self.seed, epoch])
        if self.buffer_size is None or self.buffer_size >= n:
            return rng.permutation(n)                            # buffer = whole dataset (dataset.py:51-54)
        # streaming shuffle buffer of tf.data: fill `buffer_size`, emit a random slot, refill it
        order, buf, nxt = np.empty(n, np.int64), list(range(self.buffer_size)), self.buffer_size
        for i in range(n):
            k = int(rng.integers(len(buf)))
            order[i] = buf[k]
            if nxt < n:
                buf[k] = nxt
                nxt += 1
            else:
                buf[k] = buf[-1]
                buf.pop()
        return order

    def _record_stream(self):
        """shuffle -> repeat -> filter: record indices, epochs concatenated (batches may straddle epochs)."""
        epoch = 0
        while self.num_epochs is None or epoch < self.num_epochs:
            order = self._epoch_order(epoch)
            order = order[self._keep[order]]
            if order.size == 0 and self.num_epochs is None:
                return                                           # nothing passes the filter: never spin
            for i in order:
                yield int(i)
            epoch += 1

    def _labels(self, idx):
        lab = torch.zeros((len(idx), len(self.pitches) if self.pitches else 0), dtype=torch.float32)
        for r, i in enumerate(idx):
            k = self._label_of.get(int(self.pitch[i]), -1)       # table default -1 -> all-zero row (tf.one_hot)
            if k >= 0:
                lab[r, k] = 1.0
        return lab

    def _make_batch(self, idx):
        pinned = self.device.type == "cuda"
        pcm = torch.empty((len(idx), self.waveform_length), dtype=torch.int16, pin_memory=pinned)
        decode_wav_files([self.paths[i] for i in idx], self.waveform_length, out=pcm, threads=self.threads)
        labels = self._labels(idx)
        if self.device.type != "cuda":
            return pcm, labels, None
        from . import functional as F
        with torch.cuda.stream(self._stream):
            dev = pcm.to(self.device, non_blocking=True)
            wave = F.K.pcm16_to_float(dev)
            lab = labels.pin_memory().to(self.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._stream)
        return wave, lab, (ready, pcm, dev)

    def close(self):
        """Stops the prefetch thread (it may be parked on a full queue) and ends the iteration."""
        self._done = True
        self._closed = True
        try:
            while True:
                self._queue.get_nowait()
        except queue.Empty:
            pass

    def _producer(self):
        try:
            if self.device.type == "cuda":
                torch.cuda.set_device(self.device)
            batch = []
            for i in self._records:
                batch.append(i)
                if len(batch) == self.batch_size:
                    item = self._make_batch(batch)
                    while not self._closed:
                        try:
                            self._queue.put(item, timeout=0.2)
                            break
                        except queue.Full:
                            continue
                    if self._closed:
                        return
                    batch = []
        except BaseException as e:           # surfaced by the consumer
            self._error = e
        self._queue.put(None)                # drop_remainder=True: a partial last batch is discarded

    # -------------------------------------------------------------------------------------------- consumer
    def __iter__(self):
        return self

    def __next__(self):
        if self._done:
            raise StopIteration
        item = self._queue.get()
        if item is None:
            self._done = True
            if self._error is not None:
                raise self._error
            raise StopIteration
        wave, labels, sync = item
        if sync is not None:
            torch.cuda.current_stream(self.device).wait_event(sync[0])
            wave.record_stream(torch.cuda.current_stream(self.device))
            labels.record_stream(torch.cuda.current_stream(self.device))
        return wave, labels


_PIPELINES = {}


def nsynth_input_fn(filenames, batch_size, num_epochs, shuffle, buffer_size=None, pitches=None, sources=None):
    """dataset.py:12-91.  Every call returns the next (waveforms, labels) batch; StopIteration at the end."""
    key = (tuple(filenames), batch_size, num_epochs, shuffle, buffer_size,
           tuple(pitches) if pitches is not None else None, tuple(sources) if sources is not None else None)
    pipe = _PIPELINES.get(key)
    if pipe is None:
        pipe = _PIPELINES[key] = NSynthPipeline(filenames, batch_size, num_epochs, shuffle, buffer_size, pitches, sources,
                                                device="cuda")          # the float conversion only exists on the device
    return next(pipe)


def reset_pipelines():
    """Forgets the cached pipelines (a fresh `tf.Graph()` in reference terms)."""
    for pipe in _PIPELINES.values():
        pipe.close()
    _PIPELINES.clear()
