"""Drop-in for the reference's lib/core/function.py: `train`, `fpd_train`, `validate` keep the reference
signatures (function.py:28, :99, :189) so tools/train.py / tools/fpd_train.py / tools/test.py call them unchanged.

What changes underneath:
  * fpd_train / train: when the model is the fpd_b200 hourglass and the optimizer is torch.optim.Adam, the
    batch goes through the fused path -- engine forward (+ eval-mode teacher forward without tape), ONE fused
    FPD loss+gradient kernel, the explicit backward tape -- and the gradients are handed to the caller's
    optimizer through .grad, so `optimizer.step()`, LR schedulers and checkpointing keep working. Any other
    model / criterion combination takes the generic autograd route, statement for statement like the reference.
  * per-iteration accuracy() and validate()'s flip test / decode run on the device; only [B,J] indices cross PCIe
    (the reference ships the full heat-maps to numpy every iteration: function.py:154-156, 218-240, 263).
"""
import logging
import os
import time

import numpy as np
import torch

from .evaluate import accuracy, pck_from_preds
from .inference import get_final_preds, preds_from_argmax

logger = logging.getLogger(__name__)


def _unwrap(model):
    """The fused path drives the engine of the wrapped module directly, so wrappers cannot be silently bypassed:
    nn.DataParallel over several devices would run the whole batch on one GPU -- refused (launch one process per GPU
    instead); a 1-device DataParallel / DistributedDataParallel wrapper is transparent, and
    the gradient all-reduce DDP would have done in its autograd hooks is done explicitly in _fused_step."""
    if isinstance(model, torch.nn.DataParallel) and len(model.device_ids or []) > 1:
        raise RuntimeError("fpd_b200: nn.DataParallel over %d devices is not supported by the fused path (it would use one "
                           "GPU); run one process per GPU (torch.distributed.run), device_ids=[local_rank]"
                           % len(model.device_ids))
    return model.module if hasattr(model, "module") else model


def _loss_shape_ok(config_or_none, J, h, w):
    """The fused FPD loss kernel (csrc/loss.cu) wants h*w % 64 == 0 and J <= 64; anything else takes the generic route."""
    return (h * w) % 64 == 0 and 1 <= J <= 64


_flat_grad_cache = {}


def _allreduce_grads_(net, grads):
    """One all-reduce (mean over ranks) of all parameter gradients, through one flat buffer (what DDP's bucketing does)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    n = sum(g.numel() for g in grads)
    key = (id(net), n)
    flat = _flat_grad_cache.get(key)
    if flat is None:
        _flat_grad_cache.clear()
        flat = _flat_grad_cache[key] = torch.empty(n, dtype=torch.float32, device=grads[0].device)
    views, off = [], 0
    for g in grads:
        views.append(flat[off:off + g.numel()].view(g.shape))
        off += g.numel()
    torch._foreach_copy_(views, grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / world)
    return views


def _is_fpd_net(model):
    return hasattr(_unwrap(model), "engine") and hasattr(_unwrap(model), "forward_nhwc")


def _is_fpd_criterion(c):
    from .loss import JointsMSELoss
    return isinstance(c, JointsMSELoss)


def _fused_step(model, tmodel, input, target, target_weight, alpha, use_tw):
    """student fwd -> teacher fwd -> fused loss+grad -> backward tape; returns (losses[3] dev, last output NHWC)."""
    from fpd_b200 import ops
    net = _unwrap(model)
    x = input.cuda(non_blocking=True).float().contiguous()
    eng = net.engine()
    ctx = eng.forward(x, True, record_tape=True)
    outs = [v.data for v in ctx.outs]
    t_last = None
    if tmodel is not None:
        t_last = _unwrap(tmodel).forward_nhwc(x, training=False)[-1]
    B, _, _, J = outs[0].shape
    tw = target_weight.reshape(B, J) if use_tw else torch.ones(B, J, device=x.device)
    losses, grads = ops.fpd_loss(outs, target.contiguous().float(), t_last, tw.contiguous().float(), alpha)
    pg = eng.backward(ctx, grads)
    params = [p for p in net.parameters() if p.requires_grad]      # frozen parameters keep whatever .grad they had
    glist = []
    for p in params:
        g = pg.get(p)
        glist.append(torch.zeros_like(p) if g is None else g.reshape(p.shape))
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        # one process per GPU: replicas stay identical only if every rank steps on the mean gradient
        glist = _allreduce_grads_(net, glist)
        glist = [g.clone() for g in glist]     # the flat buffer is reused next step; .grad must own its memory
    for p, g in zip(params, glist):
        p.grad = g
    return losses, outs[-1]


def _device_accuracy(out_nhwc, target_nchw):
    """PCK bookkeeping of evaluate.accuracy from device-side arg-maxes (no heat-map D2H)."""
    from fpd_b200 import ops
    B, h, w, J = out_nhwc.shape
    _, idx, mx = ops.flip_merge_argmax(out_nhwc, want_avg=False)
    tidx, tmx = ops.argmax_nchw(target_nchw.contiguous().float())
    pred, _ = preds_from_argmax(idx.cpu().numpy(), mx.cpu().numpy(), w)
    tgt, _ = preds_from_argmax(tidx.cpu().numpy(), tmx.cpu().numpy(), w)
    _, avg, cnt = pck_from_preds(pred, tgt, h, w)
    return avg, cnt, pred


def _debug_images(config, input, meta, target, pred, output, prefix):
    """The reference's debug-image hook (function.py:93-95, 184-186, 289-292; off unless config.DEBUG.DEBUG). It lives in
    the reference's own utils/vis.py (control plane, cv2 / torchvision): called when that module is importable, i.e. under
    fpd_b200.dropin.install(); `output` may be an NHWC tensor straight from the engine."""
    dbg = getattr(config, "DEBUG", None)
    if dbg is None or not getattr(dbg, "DEBUG", False):
        return
    try:
        from utils.vis import save_debug_images
    except ImportError:
        logger.warning("config.DEBUG.DEBUG is set but utils.vis is not importable: no debug images")
        return
    if output is not None and output.dim() == 4 and output.shape[1] != target.shape[1]:
        from fpd_b200 import ops
        output = ops.nhwc_to_nchw(output.contiguous())
    save_debug_images(config, input, meta, target, pred * 4, output, prefix)


def _log_train(config, writer_dict, epoch, i, n_batches, batch_time, data_time, meters, bsz):
    parts = ['Epoch: [%d][%d/%d]' % (epoch, i, n_batches),
             'Time %.3fs (%.3fs)' % (batch_time.val, batch_time.avg),
             'Speed %.1f samples/s' % (bsz / max(batch_time.val, 1e-9)),
             'Data %.3fs (%.3fs)' % (data_time.val, data_time.avg)]
    for label, m, fmt in meters:
        parts.append(('%s ' + fmt + ' (' + fmt + ')') % (label, m.val, m.avg))
    logger.info('\t'.join(parts))
    if writer_dict:
        writer = writer_dict['writer']
        step = writer_dict['train_global_steps']
        for label, m, _ in meters:
            tag = {'Loss': 'train_loss', 'POSE_Loss': 'train_pose_loss', 'KD_POSE_Loss': 'train_kd_pose_loss',
                   'Accuracy': 'train_acc'}[label]
            writer.add_scalar(tag, m.val, step)
        writer_dict['train_global_steps'] = step + 1


def train(config, train_loader, model, criterion, optimizer, epoch, output_dir, tb_log_dir, writer_dict):
    """Plain heat-map regression epoch (reference function.py:28-96): loss = sum over stacks of JointsMSELoss."""
    batch_time, data_time, losses, acc = AverageMeter(), AverageMeter(), AverageMeter(), AverageMeter()
    model.train()
    fused = _is_fpd_net(model) and _is_fpd_criterion(criterion)
    end = time.time()
    shape_checked = False
    for i, (input, target, target_weight, meta) in enumerate(train_loader):
        data_time.update(time.time() - end)
        target = target.cuda(non_blocking=True)
        target_weight = target_weight.cuda(non_blocking=True)
        if fused and not shape_checked:
            fused = _loss_shape_ok(config, target.shape[1], target.shape[2], target.shape[3])
            shape_checked = True
        if fused:
            optimizer.zero_grad()
            l3, out_nhwc = _fused_step(model, None, input, target, target_weight, 0.0, criterion.use_target_weight)
            optimizer.step()
            losses.update(float(l3[2]), input.size(0))
            avg_acc, cnt, pred = _device_accuracy(out_nhwc, target)
            output = out_nhwc
        else:
            outputs = model(input)
            if isinstance(outputs, list):
                loss = criterion(outputs[0], target, target_weight)
                for output in outputs[1:]:
                    loss += criterion(output, target, target_weight)
                output = outputs[-1]
            else:
                output = outputs
                loss = criterion(output, target, target_weight)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            losses.update(loss.item(), input.size(0))
            _, avg_acc, cnt, pred = accuracy(output.detach(), target.detach())
        acc.update(avg_acc, cnt)
        batch_time.update(time.time() - end)
        end = time.time()
        if i % config.PRINT_FREQ == 0:
            _log_train(config, writer_dict, epoch, i, len(train_loader), batch_time, data_time,
                       [('Loss', losses, '%.5f'), ('Accuracy', acc, '%.3f')], input.size(0))
            _debug_images(config, input, meta, target, pred, output.detach(), '{}_{}'.format(os.path.join(output_dir, 'train'), i))


def fpd_train(config, train_loader, model, tmodel, pose_criterion, kd_pose_criterion, optimizer, epoch,
              output_dir, tb_log_dir, writer_dict):
    """Fast-Pose-Distillation epoch (reference function.py:99-187):
    loss = (1-KD.ALPHA) * sum_s L(out_s, target) + KD.ALPHA * sum_s L(out_s, teacher_last)."""
    batch_time, data_time = AverageMeter(), AverageMeter()
    losses, pose_losses, kd_pose_losses, acc = AverageMeter(), AverageMeter(), AverageMeter(), AverageMeter()
    alpha = config.KD.ALPHA
    model.train()
    tmodel.eval()
    fused = (_is_fpd_net(model) and _is_fpd_net(tmodel) and _is_fpd_criterion(pose_criterion)
             and _is_fpd_criterion(kd_pose_criterion)
             and pose_criterion.use_target_weight == kd_pose_criterion.use_target_weight)
    end = time.time()
    shape_checked = False
    for i, (input, target, target_weight, meta) in enumerate(train_loader):
        data_time.update(time.time() - end)
        target = target.cuda(non_blocking=True)
        target_weight = target_weight.cuda(non_blocking=True)
        if fused and not shape_checked:
            fused = _loss_shape_ok(config, target.shape[1], target.shape[2], target.shape[3])
            shape_checked = True
        if fused:
            optimizer.zero_grad()
            l3, out_nhwc = _fused_step(model, tmodel, input, target, target_weight, alpha,
                                       pose_criterion.use_target_weight)
            optimizer.step()
            pose_v, kd_v, loss_v = [float(v) for v in l3.cpu()]
            avg_acc, cnt, pred = _device_accuracy(out_nhwc, target)
            last_out = out_nhwc
        else:
            outputs = model(input)
            with torch.no_grad():  # the teacher's gradients never reach the student update
                toutput = tmodel(input)
            if isinstance(toutput, list):
                toutput = toutput[-1]
            outs = outputs if isinstance(outputs, list) else [outputs]
            pose_loss = pose_criterion(outs[0], target, target_weight)
            kd_pose_loss = kd_pose_criterion(outs[0], toutput, target_weight)
            for output in outs[1:]:
                pose_loss += pose_criterion(output, target, target_weight)
                kd_pose_loss += kd_pose_criterion(output, toutput, target_weight)
            loss = (1 - alpha) * pose_loss + alpha * kd_pose_loss
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            pose_v, kd_v, loss_v = pose_loss.item(), kd_pose_loss.item(), loss.item()
            _, avg_acc, cnt, pred = accuracy(outs[-1].detach(), target.detach())
            last_out = outs[-1].detach()
        pose_losses.update(pose_v, input.size(0))
        kd_pose_losses.update(kd_v, input.size(0))
        losses.update(loss_v, input.size(0))
        acc.update(avg_acc, cnt)
        batch_time.update(time.time() - end)
        end = time.time()
        if i % config.PRINT_FREQ == 0:
            _log_train(config, writer_dict, epoch, i, len(train_loader), batch_time, data_time,
                       [('POSE_Loss', pose_losses, '%.5f'), ('KD_POSE_Loss', kd_pose_losses, '%.5f'),
                        ('Loss', losses, '%.5f'), ('Accuracy', acc, '%.3f')], input.size(0))
            _debug_images(config, input, meta, target, pred, last_out, '{}_{}'.format(os.path.join(output_dir, 'train'), i))


def validate(config, val_loader, val_dataset, model, criterion, output_dir, tb_log_dir, writer_dict=None):
    """Evaluation epoch (reference function.py:189-333) with the flip test, shift, average, arg-max on the device."""
    from fpd_b200 import ops
    batch_time, losses, acc = AverageMeter(), AverageMeter(), AverageMeter()
    model.eval()
    num_samples = len(val_dataset)
    all_preds = np.zeros((num_samples, config.MODEL.NUM_JOINTS, 3), dtype=np.float32)
    all_boxes = np.zeros((num_samples, 6))
    image_path, filenames, imgnums = [], [], []
    idx = 0
    fpd_net = _is_fpd_net(model)
    perm_cache = {}
    with torch.no_grad():
        end = time.time()
        for i, (input, target, target_weight, meta) in enumerate(val_loader):
            x = input.cuda(non_blocking=True).float().contiguous()
            if fpd_net:
                net = _unwrap(model)
                hm = net.forward_nhwc(x, training=False)[-1]
                hm_f, perm = None, None
                if config.TEST.FLIP_TEST:
                    hm_f = net.forward_nhwc(x.flip(3).contiguous(), training=False)[-1]
                    J = hm.shape[-1]
                    if J not in perm_cache:
                        p = list(range(J))
                        for a, b in val_dataset.flip_pairs:
                            p[a], p[b] = b, a
                        perm_cache[J] = torch.tensor(p, dtype=torch.int32, device=x.device)
                    perm = perm_cache[J]
                avg_nhwc, _, _ = ops.flip_merge_argmax(hm, hm_f, perm, shift=bool(config.TEST.SHIFT_HEATMAP))
                output = ops.nhwc_to_nchw(avg_nhwc)
            else:
                outputs = model(x)
                output = outputs[-1] if isinstance(outputs, list) else outputs
                if config.TEST.FLIP_TEST:
                    from utils.transforms import flip_back  # resolved like the reference (lib on sys.path)
                    of = model(x.flip(3))
                    of = of[-1] if isinstance(of, list) else of
                    of = flip_back(of, val_dataset.flip_pairs).clone()
                    if config.TEST.SHIFT_HEATMAP:
                        of[:, :, :, 1:] = of.clone()[:, :, :, 0:-1]
                    output = (output + of) * 0.5
            target = target.cuda(non_blocking=True)
            target_weight = target_weight.cuda(non_blocking=True)
            loss = criterion(output, target, target_weight)
            num_images = input.size(0)
            losses.update(loss.item(), num_images)
            _, avg_acc, cnt, pred = accuracy(output, target)
            acc.update(avg_acc, cnt)
            batch_time.update(time.time() - end)
            end = time.time()
            c = meta['center'].numpy()
            s = meta['scale'].numpy()
            score = meta['score'].numpy()
            preds, maxvals = get_final_preds(config, output, c, s)
            all_preds[idx:idx + num_images, :, 0:2] = preds[:, :, 0:2]
            all_preds[idx:idx + num_images, :, 2:3] = maxvals
            all_boxes[idx:idx + num_images, 0:2] = c[:, 0:2]
            all_boxes[idx:idx + num_images, 2:4] = s[:, 0:2]
            all_boxes[idx:idx + num_images, 4] = np.prod(s * 200, 1)
            all_boxes[idx:idx + num_images, 5] = score
            image_path.extend(meta['image'])
            idx += num_images
            if i % config.PRINT_FREQ == 0:
                logger.info('Test: [%d/%d]\tTime %.3f (%.3f)\tLoss %.4f (%.4f)\tAccuracy %.3f (%.3f)' % (
                    i, len(val_loader), batch_time.val, batch_time.avg, losses.val, losses.avg, acc.val, acc.avg))
                _debug_images(config, input, meta, target, pred, output, '{}_{}'.format(os.path.join(output_dir, 'val'), i))
        name_values, perf_indicator = val_dataset.evaluate(config, all_preds, output_dir, all_boxes, image_path,
                                                           filenames, imgnums)
        model_name = config.MODEL.NAME
        for nv in (name_values if isinstance(name_values, list) else [name_values]):
            _print_name_value(nv, model_name)
        if writer_dict:
            writer = writer_dict['writer']
            step = writer_dict['valid_global_steps']
            writer.add_scalar('valid_loss', losses.avg, step)
            writer.add_scalar('valid_acc', acc.avg, step)
            for nv in (name_values if isinstance(name_values, list) else [name_values]):
                writer.add_scalars('valid', dict(nv), step)
            writer_dict['valid_global_steps'] = step + 1
    return perf_indicator


def _print_name_value(name_value, full_arch_name):
    names, values = list(name_value.keys()), list(name_value.values())
    logger.info('| Arch ' + ' '.join('| %s' % n for n in names) + ' |')
    logger.info('|---' * (len(names) + 1) + '|')
    if len(full_arch_name) > 15:
        full_arch_name = full_arch_name[:8] + '...'
    logger.info('| ' + full_arch_name + ' ' + ' '.join('| %.3f' % v for v in values) + ' |')


class AverageMeter(object):
    """Running value / average (same fields the reference's meter exposes: val, avg, sum, count)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count if self.count != 0 else 0
