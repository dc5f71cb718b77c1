"""train() / validate() of the reference's torch_train.py:159-305 with the per-batch host round trips removed.

The reference synchronises the host several times per batch: `batch_loss.cpu().item()`, one `.cpu().item()` per metric,
and (with a summary writer) an abs-max scan of every gradient (torch_train.py:198-213, 268-281).  Here one fused device
reduction per batch (snb_loss_iou_reduce) yields the loss AND every reference metric (JaccardScore, PixelAccuracy) as
device scalars that are written into a preallocated [n_batches, 1 + n_metrics] table; the AverageMeters are filled from
that table with ONE device -> host copy per epoch.  In train mode models that offer `train_step` (LinkNet34) run forward,
loss, its gradient and the whole backward pass as two CUDA-graph replays without autograd.  Same signatures and return
values as the reference: (losses: AverageMeter, scores: {name: AverageMeter}).  TensorBoard logging is out of scope
(SURVEY 2); `summary_writer`, when given, receives the per-epoch scalars only.
"""
import torch

from .lib import losses as L
from .lib import metrics as M
from .lib.train_utils import AverageMeter, PRCurveMeter


def _batch_scalars(loss, metrics, outputs, y, row):
    """loss + metrics of one batch into `row` (device float tensor [1 + n_metrics]); one fused reduction when possible."""
    fused = hasattr(loss, 'fused_coefficients') and getattr(loss, 'reduce', True) is not False and all(
        isinstance(m, (M.JaccardScore, M.PixelAccuracy)) for m in metrics.values())
    if fused:
        k = dict(c_bce=0.0, c_focal=0.0, gamma=0.0, c_jac=0.0, smooth_num=0.0, smooth_den=0.0)
        k.update(loss.fused_coefficients(outputs.numel()))
        sums, counts = L.fused_sums(outputs, y, focal_gamma=k['gamma'] if k['c_focal'] else None)
        row[0] = L._combine(sums, k['c_bce'], k['c_focal'], k['gamma'], k['c_jac'], k['smooth_num'], k['smooth_den'])
        for j, m in enumerate(metrics.values()):
            if isinstance(m, M.JaccardScore):
                row[1 + j] = (sums[1] / (sums[2] + sums[3] - sums[1] + 1e-7)).float()
            else:
                row[1 + j] = (counts[0] + counts[3]).float() / y.numel()
        return
    row[0] = loss(outputs, y).detach().float()
    for j, m in enumerate(metrics.values()):
        row[1 + j] = torch.as_tensor(m(outputs, y), device=row.device).detach().float()


def _fill_meters(table, metrics):
    host = table.cpu().tolist()                      # the epoch's only device -> host copy of scalars
    losses = AverageMeter()
    scores = {key: AverageMeter() for key in metrics}
    for r in host:
        losses.update(r[0])
        for j, key in enumerate(metrics):
            scores[key].update(r[1 + j])
    return losses, scores


def train(model, loss, optimizer, dataloader, epoch: int, metrics={}, summary_writer=None):
    """torch_train.py:159-237."""
    batches = list(dataloader) if not hasattr(dataloader, '__len__') else dataloader
    n_batches = len(batches)
    table = None
    with torch.set_grad_enabled(True):
        model.train()
        for batch_index, (x, y) in enumerate(batches):
            x, y = x.cuda(non_blocking=True), y.cuda(non_blocking=True)
            if table is None:
                table = torch.zeros((n_batches, 1 + len(metrics)), dtype=torch.float32, device=x.device)
            optimizer.zero_grad()
            if hasattr(model, 'train_step') and hasattr(loss, 'fused_coefficients'):
                _, outputs = model.train_step(x, y, loss)                     # forward + loss gradient + backward, no autograd
            else:
                outputs = model(x)
                (x.size(0) * loss(outputs, y)).backward()
            optimizer.step()
            with torch.no_grad():
                _batch_scalars(loss, metrics, outputs.detach(), y, table[batch_index])
    losses, scores = _fill_meters(table, metrics) if table is not None else (AverageMeter(), {k: AverageMeter() for k in metrics})
    if summary_writer is not None:
        summary_writer.add_scalar('train/epoch/loss', losses.avg, epoch)
        for key, value in scores.items():
            summary_writer.add_scalar('train/epoch/' + key, value.avg, epoch)
    return losses, scores


def validate(model, loss, dataloader, epoch: int, metrics=dict(), summary_writer=None, pr_meter=None):
    """torch_train.py:240-305; the PR curve of the last batch (as the reference computes it) lands in `pr_meter`."""
    batches = list(dataloader) if not hasattr(dataloader, '__len__') else dataloader
    n_batches = len(batches)
    table, outputs, y = None, None, None
    with torch.set_grad_enabled(False):
        model.eval()
        for batch_index, (x, y) in enumerate(batches):
            x, y = x.cuda(non_blocking=True), y.cuda(non_blocking=True)
            if table is None:
                table = torch.zeros((n_batches, 1 + len(metrics)), dtype=torch.float32, device=x.device)
            outputs = model(x)
            _batch_scalars(loss, metrics, outputs, y, table[batch_index])
        if outputs is not None:
            pr_meter = PRCurveMeter() if pr_meter is None else pr_meter
            pr_meter.update(outputs, y)
    losses, scores = _fill_meters(table, metrics) if table is not None else (AverageMeter(), {k: AverageMeter() for k in metrics})
    if summary_writer is not None:
        summary_writer.add_scalar('val/epoch/loss', losses.avg, epoch)
        for key, value in scores.items():
            summary_writer.add_scalar('val/epoch/' + key, value.avg, epoch)
    return losses, scores


def get_loss(loss):
    """torch_train.py:75-96: the registry strings of the binary losses."""
    loss = str.lower(loss)
    if loss == 'smooth_jaccard':
        return L.SmoothJaccardLoss()
    if loss == 'jaccard':
        return L.JaccardLoss()
    if loss == 'bce_jaccard':
        return L.BCEWithLogitsLossAndSmoothJaccard()
    if loss == 'focal':
        return L.FocalLossBinary(size_average=False)
    if loss == 'bce':
        return L.BCEWithSigmoidLoss()
    raise ValueError(loss)


def get_model(model_name, patch_size=None, num_channels=3):
    """torch_train.py:99-147 for the models built on the native engine; the ResNet101/152-based registry entries (gcn,
    psp_net, duc, linknext, dilated_linknet34, squeezenet) need dilation / bilinear up-sampling kernels and are out of scope."""
    from .lib import models as MM

    model_name = str.lower(model_name)
    if num_channels != 3:
        raise NotImplementedError("the native first-layer kernels take 3-channel images")
    table = {'unet': MM.UNet, 'unet_abn': MM.UNetABN, 'zf_unet': MM.ZF_UNET,
             'unet11': lambda: MM.UNet11(pretrained=True), 'unet16': lambda: MM.UNet16(pretrained=True),
             'linknet34': lambda: MM.LinkNet34(pretrained=True, num_channels=num_channels, num_classes=1),
             'tiramisu67': lambda: MM.FCDenseNet67(n_classes=1)}
    if model_name in table:
        return table[model_name]()
    if model_name in ('dilated_linknet34', 'linknext', 'gcn', 'gcn34', 'psp_net', 'duc', 'duc_dc', 'squeezenet'):
        raise NotImplementedError("registry model %r is outside the hot-path scope (SURVEY 8f.4)" % model_name)
    raise ValueError(model_name)
