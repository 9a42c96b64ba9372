"""Input stage of the training / evaluation loops (the step in front of the hot path).

The reference decodes and normalises every image on the loader workers (``transforms.Resize`` -> ``ToTensor`` ->
``Normalize((0.5,)*3, (0.5,)*3)``, train_3_encoder.py:231-237), collates fp32 batches, round-trips the reference images through
``.cpu().numpy()`` (``Data_Loading``, dataset.py:377-378,389-390) and copies everything to the device with blocking
``.to(device)`` calls (:401) at the top of every iteration.  At thousands of images per second that stage is what the GPU
waits for.  Here:

* ``u8_transform(size)``      what the dataset workers run instead: resize + uint8 HWC array (no float conversion on the host);
* ``DevicePrefetcher``        wraps the loader iterator: pins each batch, issues its H2D copy on a side stream one batch ahead
                              of the consumer, and normalises uint8 batches on the device (``fm_im2tensor_f32``, bit-exact with
                              ToTensor + Normalize) -- 1/4 of the PCIe bytes, no host float pass, no blocking copy;
* ``Data_Loading``            same name, arguments and return values as dataset.py:361-413, fed by prefetchers (or by the
                              reference's loaders): batches that are already on the device stay there, the swapped reference batch
                              is an index on the device instead of a numpy round trip.
"""
import numpy as np
import torch

from . import ops


def u8_transform(size):
    """PIL image -> uint8 tensor [H,W,3] (resized like ``transforms.Resize(size)``); the float conversion and the
    normalisation happen on the device."""
    from torchvision import transforms

    resize = transforms.Resize(size)

    def f(img):
        return torch.from_numpy(np.asarray(resize(img.convert("RGB")), dtype=np.uint8).copy())
    return f


class DevicePrefetcher:
    """Iterator over device-resident batches, one H2D copy in flight ahead of the consumer.

    ``it`` yields a tensor or a tuple / list of tensors (what the reference's ``sample_data(loader)`` yields,
    train_3_encoder.py:203-207).  uint8 ``[B,H,W,3]`` members are normalised to fp32 ``[B,3,H,W]`` in [-1,1] on the device;
    float members are copied as they are."""

    def __init__(self, it, device, mean=0.5, std=0.5):
        self.it = iter(it)
        self.device = torch.device(device)
        self.mean, self.std = mean, std
        self.stream = torch.cuda.Stream(self.device)
        self._next = None
        self._preload()

    def _to_device(self, t):
        if not torch.is_tensor(t):
            return t
        if not t.is_cuda:
            t = t.pin_memory() if not t.is_pinned() else t
            t = t.to(self.device, non_blocking=True)
        if t.dtype == torch.uint8 and t.ndim == 4 and t.shape[-1] == 3:
            t = ops.im2tensor_batch(t, self.mean, self.std)
        return t

    def _preload(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            if isinstance(batch, (tuple, list)):
                self._next = type(batch)(self._to_device(t) for t in batch)
            else:
                self._next = self._to_device(batch)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)
        batch = self._next
        for t in (batch if isinstance(batch, (tuple, list)) else (batch,)):
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(cur)
        self._preload()
        return batch


def _swap_pairs(n, device):
    """Index of ``Swap_List_Pair(range(n))`` (dataset.py:342-358): every even / odd neighbour pair exchanged."""
    idx = torch.arange(n, device=device)
    return idx + 1 - 2 * (idx % 2)


def Data_Loading(rec_loader, ds_loader, ds_flag, device, extreme_loader=None, extreme_ds_flag=False,
                 pure_ffhq_loader=None, ds_dataset_type=None):
    """dataset.py:361-413 with the same contract: returns ``(g_input, r_input, g_ref)`` on ``device`` (five tensors for
    ``ds_dataset_type == 'FFHQ'``).  No ``.cpu().numpy()`` round trip: the reference batch is a device-side copy / index of
    the generator input; inputs already on the device (``DevicePrefetcher``) are not copied again."""
    dev = torch.device(device)

    def put(t):
        return t if t.is_cuda else t.to(dev, non_blocking=True)
    if ds_dataset_type is None:
        if ds_flag is False:
            g_input, r_input = next(rec_loader)
            g_input, r_input = put(g_input), put(r_input)
            g_ref = g_input.clone()
        else:
            g_input, r_input = next(extreme_loader if extreme_ds_flag else ds_loader)
            g_input, r_input = put(g_input), put(r_input)
            swap = _swap_pairs(g_input.shape[0], g_input.device)
            r_input = r_input[swap]
            g_ref = g_input[swap]
            if extreme_ds_flag:
                even = torch.arange(g_input.shape[0] // 2, device=g_input.device) * 2     # only the even indices
                g_input, r_input, g_ref = g_input[even], r_input[even], g_ref[even]
        return g_input, r_input, g_ref
    if ds_dataset_type == 'FFHQ':
        ffhq_ref = put(next(pure_ffhq_loader))
        g_input, r_input, r_edit_input = (put(t) for t in next(ds_loader))
        return g_input, r_input, r_edit_input, g_input.clone(), ffhq_ref
    raise ValueError(f"unknown ds_dataset_type {ds_dataset_type!r}")
