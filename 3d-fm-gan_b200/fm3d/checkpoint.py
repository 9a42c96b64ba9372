"""Checkpoint output that does not stop the training loop.

The reference writes its checkpoint with one blocking ``torch.save`` of ten state dicts straight from device memory
(train_3_encoder.py:735-753: G, g_ema, D, three encoders, two Adam states -- ~3.5 GB at 256x256): the step stalls for the
device->host copy, the pickling and the disk write.  ``AsyncCheckpointWriter.save(obj, path)`` snapshots every CUDA tensor
of ``obj`` into pinned host buffers with non-blocking copies on a side stream (ordered after the work already queued on
the calling stream, so the snapshot is consistent), returns at once, and a background thread waits for the copies and runs
``torch.save`` on the host-side structure.  The file is byte-for-byte what ``torch.save`` of the CPU state dicts writes, so
``torch.load`` / ``Module_To_Train_Setup`` (train_3_encoder.py:330-347) read it unchanged.
"""
import os
import threading

import torch


class AsyncCheckpointWriter:
    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else None
        self._stream = None
        self._thread = None
        self._error = None
        self._pinned = {}            # (shape, dtype) -> reusable pinned buffers of the previous snapshot

    def _snapshot(self, obj, pool):
        if torch.is_tensor(obj):
            if not obj.is_cuda:
                return obj.detach().clone()
            key = (tuple(obj.shape), obj.dtype)
            free = pool.setdefault(key, [])
            buf = free.pop() if free else torch.empty(obj.shape, dtype=obj.dtype, pin_memory=True)
            buf.copy_(obj.detach(), non_blocking=True)
            self._used.setdefault(key, []).append(buf)
            return buf
        if isinstance(obj, dict):
            return type(obj)((k, self._snapshot(v, pool)) for k, v in obj.items())
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._snapshot(v, pool) for v in obj)
        return obj

    def save(self, obj, path):
        """Snapshot ``obj`` (nested dicts / lists of tensors and plain values) and write it to ``path`` in the background.
        A previous save still in flight is waited for first (its pinned buffers are reused)."""
        self.wait()
        if torch.cuda.is_available():
            dev = self.device or torch.device("cuda", torch.cuda.current_device())
            if self._stream is None:
                self._stream = torch.cuda.Stream(dev)
            self._stream.wait_stream(torch.cuda.current_stream(dev))
            self._used = {}
            with torch.cuda.stream(self._stream):
                host = self._snapshot(obj, self._pinned)
            done = torch.cuda.Event()
            done.record(self._stream)
        else:
            self._used = {}
            host, done = self._snapshot(obj, self._pinned), None
        used = self._used

        def work():
            try:
                if done is not None:
                    done.synchronize()
                tmp = path + ".tmp"
                # plain (unpinned) copies would double the host memory; torch.save reads the pinned buffers directly
                torch.save(host, tmp)
                os.replace(tmp, path)
            except Exception as e:      # noqa: BLE001  (reported by wait())
                self._error = e
            finally:
                for k, bufs in used.items():
                    self._pinned.setdefault(k, []).extend(bufs)
        self._thread = threading.Thread(target=work, daemon=True)
        self._thread.start()

    def wait(self):
        """Block until the last save has reached the disk; re-raises its error, if any."""
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self._error is not None:
            e, self._error = self._error, None
            raise e
