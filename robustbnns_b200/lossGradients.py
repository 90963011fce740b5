"""Drop-in for the reference's `lossGradients.py` hot path (lossGradients.py:20-76).

`loss_gradient` / `loss_gradients` keep the reference's signatures, return types,
prints and pickle side effect; the per-image x per-sample Python loop
(lossGradients.py:29-38, :56-60) is replaced by one batched, sample-major
evaluation on the GPU: sum_s dCE(softmax(softmax(f_s(x))), y)/dx over the bank rows
of seeds 0..S-1, one allreduce of the [B, D] partial sums across ranks, times 1/S.
"""
import numpy as np
import torch

from . import dist as rdist
from ._lib import HEAD_MEAN_OF_GRADS
from .savedir import DATA
from .utils import load_from_pickle, save_to_pickle

DEBUG = False
MAX_BATCH = 16384     # inputs per device pass


_COPY_STREAMS = {}


def _h2d(t, device, dtype):
    """Host -> device.  A pinned host tensor of the right dtype is copied on a side stream, so that whatever the caller
    already enqueued on the compute stream (the K-sample of the posterior draws) runs while the bytes cross PCIe; the
    compute stream waits for the copy before it continues."""
    if t.device.type != "cpu" or t.dtype != dtype or not t.is_pinned() or not torch.cuda.is_available():
        return t.to(device=device, dtype=dtype)
    device = torch.device(device)
    cs = _COPY_STREAMS.get(device.index)
    if cs is None:
        cs = _COPY_STREAMS[device.index] = torch.cuda.Stream(device)
    main = torch.cuda.current_stream(device)
    with torch.cuda.stream(cs):
        d = t.to(device, non_blocking=True)
    main.wait_stream(cs)
    d.record_stream(main)
    return d


def _to_device(eng, images, labels):
    """Inputs and labels on the engine's device.  Host tensors under torch.distributed: every rank copies only its block
    of rows over PCIe and the blocks are all-gathered over NVLink (dist.gather_rows_from_host)."""
    images, labels = torch.as_tensor(images), torch.as_tensor(labels)
    if images.device.type == "cpu" and rdist.world()[1] > 1:
        images = rdist.gather_rows_from_host(images, eng.device, torch.float32)
    else:
        images = _h2d(images, eng.device, torch.float32)
    if labels.device.type == "cpu" and rdist.world()[1] > 1:
        labels = rdist.gather_rows_from_host(labels.to(torch.int32), eng.device, torch.int32)
    else:
        labels = _h2d(labels, eng.device, torch.int32)
    return images, labels


def expected_loss_gradients_block(net, images, labels, n_samples):
    """Row-sharded result for multi-GPU callers: like `expected_loss_gradients`, but the [B, D] partial sums of the ranks
    are reduce-SCATTERed, so every rank ends up with (and only has to read back) its own block of rows.
    Returns (block [hi - lo, *input_shape] on the device, (lo, hi)); one process: the whole batch, (0, B)."""
    eng = net.engine()
    net._rows(n_samples, list(range(n_samples)))                # K-sample first (cached afterwards): it runs while the inputs cross PCIe
    images, labels = _to_device(eng, images, labels)
    B = images.shape[0]
    rank, world = rdist.world()
    if world == 1 or B > MAX_BATCH:
        g = expected_loss_gradients(net, images, labels, n_samples)
        lo, hi, _ = rdist.row_block(B, rank, world)
        return g[lo:hi], (lo, hi)
    lo, hi, per = rdist.row_block(B, rank, world)
    rows, _ = net._rows(n_samples, list(range(n_samples)))
    x = images.contiguous()
    buf = torch.zeros((per * world, eng.D), dtype=torch.float32, device=eng.device) if per * world != B else \
        torch.empty((B, eng.D), dtype=torch.float32, device=eng.device)
    eng.input_grad_sum(HEAD_MEAN_OF_GRADS, x, labels.contiguous(), rows[0], rows[1], out=buf[:B])
    blk = rdist.reduce_scatter_rows_(buf, per)[:hi - lo]
    return (blk * (1.0 / float(n_samples))).reshape((hi - lo,) + tuple(x.shape[1:])), (lo, hi)


def expected_loss_gradients(net, images, labels, n_samples):
    """[B, *input_shape] tensor on the device: (1/S) sum_{s<S} dL_s/dx for a batch
    (`labels` are class indices).  Sample s is seed s (lossGradients.py:33)."""
    eng = net.engine()
    rows, _ = net._rows(n_samples, list(range(n_samples)))      # K-sample first: it runs while the inputs cross PCIe
    images, labels = _to_device(eng, images, labels)
    outs = []
    for b0 in range(0, images.shape[0], MAX_BATCH):
        x = images[b0:b0 + MAX_BATCH].contiguous()
        g = eng.input_grad_sum(HEAD_MEAN_OF_GRADS, x, labels[b0:b0 + MAX_BATCH].contiguous(), rows[0], rows[1])
        rdist.allreduce_sum_(g)
        outs.append((g * (1.0 / float(n_samples))).reshape(x.shape))
    return torch.cat(outs) if len(outs) != 1 else outs[0]


def expected_loss_gradients_prefix(net, images, labels, n_samples_list):
    """The expected loss gradients for EVERY sample count of `n_samples_list` from one pass over max(n_samples_list)
    posterior samples.  Sample s is seed s (lossGradients.py:33), so the run with n samples uses a prefix of the run with
    m > n samples: the reference's main() (lossGradients.py:132-151) and plot_gradients_components._get_gradients
    (:125-142) evaluate [1, 10, 50, 100] one after the other -- 161 sample evaluations per image for 100 distinct ones.
    Here the sample range is cut at the list's sizes, each piece is evaluated once and the running sum is divided by the
    prefix length.  Returns a list of [B, *input_shape] device tensors in the order of `n_samples_list`."""
    sizes = [int(n) for n in n_samples_list]
    if not sizes or min(sizes) < 1:
        raise ValueError("n_samples_list must hold positive sample counts")
    order = sorted(set(sizes))
    eng = net.engine()
    images, labels = _to_device(eng, images, labels)
    rows, _ = net._rows(order[-1], list(range(order[-1])))       # every sample placed (drawn / uploaded) exactly once
    rank, world = rdist.world()
    # rank r holds samples r, r + W, ... in consecutive rows: the prefix of n samples is its first local_count(n) rows
    cuts = [rows[0] + rdist.local_count(n, rank, world) for n in order]
    outs = dict((n, []) for n in order)
    for b0 in range(0, images.shape[0], MAX_BATCH):
        x = images[b0:b0 + MAX_BATCH].contiguous()
        y = labels[b0:b0 + MAX_BATCH].contiguous()
        acc, lo = None, rows[0]
        for n, hi in zip(order, cuts):
            g = eng.input_grad_sum(HEAD_MEAN_OF_GRADS, x, y, lo, hi)
            rdist.allreduce_sum_(g)
            acc = g.clone() if acc is None else acc + g
            outs[n].append((acc * (1.0 / float(n))).reshape(x.shape))
            lo = hi
    return [torch.cat(outs[n]) if len(outs[n]) != 1 else outs[n][0] for n in sizes]


def loss_gradient(net, image, label, n_samples=None):
    """One image [ch,h,w] + one-hot label -> expected loss gradient [ch,h,w] (lossGradients.py:20-50)."""
    if not n_samples:
        # the reference's deterministic branch dereferences undefined names (lossGradients.py:43)
        raise NameError("name 'net_copy' is not defined")
    image = torch.as_tensor(image).unsqueeze(0)
    label = torch.as_tensor(label).argmax(-1).unsqueeze(0)
    return expected_loss_gradients(net, image, label, n_samples)[0]


def loss_gradients(net, data_loader, device, filename, savedir, n_samples=None):
    print(f"\n === Loss gradients on {len(data_loader.dataset)} input images:")
    images, labels = [], []
    for x_batch, y_batch in data_loader:
        images.append(torch.as_tensor(x_batch))
        labels.append(torch.as_tensor(y_batch).argmax(-1))
    images, labels = torch.cat(images), torch.cat(labels)
    if not n_samples:
        raise NameError("name 'net_copy' is not defined")
    grads = expected_loss_gradients(net, images, labels, n_samples)
    print(f"\nmin = {grads.min():.4f} \t max = {grads.max():.4f}")
    grads = grads.cpu().detach().numpy().squeeze()
    rank, world = rdist.real_world()
    if rank == 0:                                   # one writer: concurrent open(..., "wb") by every rank truncates the file
        save_loss_gradients(grads, n_samples, filename, savedir)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return grads


def loss_gradients_list(net, data_loader, device, filename, savedir, n_samples_list):
    """`loss_gradients` for every entry of `n_samples_list` in ONE pass over the posterior samples
    (`expected_loss_gradients_prefix`): the same prints, the same pickle per sample count
    (`<filename>_samp=<n>_lossGrads.pkl`) and the same numpy arrays the reference's loop over
    `posterior_samples_list` produces (lossGradients.py:148-151, plot_gradients_components.py:129-140)."""
    images, labels = [], []
    for x_batch, y_batch in data_loader:
        images.append(torch.as_tensor(x_batch))
        labels.append(torch.as_tensor(y_batch).argmax(-1))
    images, labels = torch.cat(images), torch.cat(labels)
    grads_list = expected_loss_gradients_prefix(net, images, labels, n_samples_list)
    rank, world = rdist.real_world()
    out = []
    for n_samples, grads in zip(n_samples_list, grads_list):
        print(f"\n === Loss gradients on {len(data_loader.dataset)} input images:")
        print(f"\nmin = {grads.min():.4f} \t max = {grads.max():.4f}")
        grads = grads.cpu().detach().numpy().squeeze()
        if rank == 0:
            save_loss_gradients(grads, n_samples, filename, savedir)
        out.append(grads)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return out


def save_loss_gradients(loss_gradients, n_samples, filename, savedir, relpath=DATA):
    save_to_pickle(data=loss_gradients, path=relpath + savedir,
                   filename=filename + "_samp=" + str(n_samples) + "_lossGrads.pkl")


def load_loss_gradients(n_samples, filename, savedir, relpath=DATA):
    path = relpath + savedir + filename + "_samp=" + str(n_samples) + "_lossGrads.pkl"
    return load_from_pickle(path=path)


def compute_vanishing_norms_idxs(loss_gradients, n_samples_list, norm):
    """Indices of images whose gradient norm is non-increasing along `n_samples_list`
    (lossGradients.py:78-127), vectorised; prints the reference's three summary lines."""
    loss_gradients = np.asarray(loss_gradients)
    if loss_gradients.shape[1] != len(n_samples_list):
        raise ValueError("Second dimension should equal the length of `n_samples_list`")
    flat = loss_gradients.reshape(loss_gradients.shape[0], loss_gradients.shape[1], -1)
    if norm == "linfty":
        norms = np.abs(flat).max(-1)
    elif norm == "l2":
        norms = np.sqrt((flat.astype(np.float64) ** 2).sum(-1)).astype(flat.dtype)
    else:
        raise ValueError("norm must be 'linfty' or 'l2'")
    nonnull = norms[:, 0] != 0.0
    # running minimum semantics of the reference: a step counts when it does not exceed the last accepted norm
    cur = norms[:, 0].copy()
    count = np.zeros(len(norms), dtype=np.int64)
    for j in range(norms.shape[1]):
        ok = norms[:, j] <= cur
        cur = np.where(ok, norms[:, j], cur)
        count += ok
    vanishing = nonnull & (count == norms.shape[1])
    idxs = [int(i) for i in np.nonzero(vanishing)[0]]
    n = len(loss_gradients)
    print(f"vanishing gradients = {vanishing.sum()/n} %")
    print(f"increasing gradients = {(nonnull & ~vanishing).sum()/n} %")
    print(f"null gradients = {(~nonnull).sum()/n} %")
    print("\nvanishing_gradients_idxs = ", idxs)
    return idxs
