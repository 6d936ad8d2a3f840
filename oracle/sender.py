"""CPU restatement of the reference's sender loop around the sampling path -- TEST INFRASTRUCTURE ONLY (imported by
tests/ alone; the product is evcdiff/sender.py).

Follows /root/reference/city_sender.py:
  cal_psnr            :255-258
  decide_5to5         :353-374   (PSNR accept decision)
  decide_5to5_lpips   :376-406   (the same prefix rule on a perceptual distance, accepted while <= threshold; the LPIPS
                                  network -- lpips==0.1.4 with AlexNet weights, requirements.txt:66, not in this image -- is
                                  an injected callable)
  update              :408-437   (SenderCity.update)
  encode_video        :519-550   (the `while x_ge.shape[1] < 30` loop of the driver script, one video at a time)

The script itself cannot be imported (top-level argparse, compressai, I3D weights), so parity is pinned on the call
sites: numpy float64 PSNR, accept the longest prefix of predicted frames whose PSNR >= threshold, append them with flag
0; when nothing was accepted, code two more keyframes with flag 1.  `generate_frame` (the diffusion sampler) and
`compress` (the ELIC keyframe codec) are injected callables.
"""
import numpy as np
import torch


def cal_psnr(img1, img2, maxvalue=1.0):
    """city_sender.py:255-258."""
    img1, img2 = img1.astype(np.float64), img2.astype(np.float64)
    mse = np.mean((img1 - img2) ** 2)
    return 10 * np.log10((maxvalue ** 2) / mse)


def decide_5to5(pred, gt, threshold):
    """city_sender.py:353-374 for one video (batchsize 1, as the script runs it).  pred (1,5,C,H,W), gt (1,F<=5,C,H,W)
    numpy.  Returns (new_d (1,n), new_ge (1,n,C,H,W)) with n = length of the accepted prefix."""
    batchsize, frames_num = gt.shape[0], gt.shape[1]
    new_d, new_ge = [], []
    for i in range(batchsize):
        for j in range(frames_num):
            if cal_psnr(pred[i][j], gt[i][j]) >= threshold:
                new_ge.append(pred[i][j])
                new_d.append(0)
            else:
                break
    new_d = np.array(new_d, dtype=np.int64).reshape(batchsize, -1)
    new_ge = np.array(new_ge, dtype=pred.dtype).reshape((batchsize, -1) + tuple(gt.shape[2:]))
    return new_d, new_ge


def decide_5to5_lpips(pred, gt, threshold, loss_fn):
    """city_sender.py:376-406 for one video: accept predicted frames while loss_fn(pred_j, gt_j) <= threshold.
    pred, gt torch tensors (1,5,C,H,W) / (1,F,C,H,W); loss_fn stands for `self.loss_fn_alex` (city_sender.py:302,389)."""
    batchsize, frames_num = gt.shape[0], gt.shape[1]
    new_d, new_ge = [], []
    for i in range(batchsize):
        for j in range(frames_num):
            if float(loss_fn(pred[i][j].float(), gt[i][j].float())) <= threshold:
                new_ge.append(pred[i][j].detach().cpu().numpy())
                new_d.append(0)
            else:
                break
    new_d = np.array(new_d, dtype=np.int64).reshape(batchsize, -1)
    new_ge = np.array(new_ge, dtype=np.float32).reshape((batchsize, -1) + tuple(gt.shape[2:]))
    return new_d, new_ge


def update(x_gt, x_ge, d, generate_frame, threshold, num_cond=2, num_pred=5, lpips_fn=None):
    """city_sender.py:408-437 (use_psnr branch, or use_lpips when `lpips_fn` is given).  x_gt (1,T,C,H,W), x_ge (1,t,C,H,W)
    torch tensors in [0,1]; generate_frame maps the last `num_cond` reconstructed frames (1, num_cond*C, H, W) to
    (1, num_pred, C, H, W)."""
    B, T, C, H, W = x_ge.shape
    idx = x_ge.shape[1]
    frames_gt = x_gt[:, idx:idx + num_pred]
    input_frames = x_ge[:, -num_cond:].reshape(B, -1, H, W)
    pred = generate_frame(input_frames).reshape(B, -1, C, H, W)
    if lpips_fn is not None:
        new_d, new_ge = decide_5to5_lpips(torch.as_tensor(pred), frames_gt, threshold, lpips_fn)
    else:
        new_d, new_ge = decide_5to5(np.asarray(pred.cpu().numpy() if torch.is_tensor(pred) else pred), frames_gt.numpy(), threshold)
    d = np.concatenate((d, new_d), axis=1)
    x_ge = torch.from_numpy(np.concatenate((x_ge.numpy(), new_ge.astype(x_ge.numpy().dtype)), axis=1))
    return d, x_ge


def encode_video(x_gt, generate_frame, threshold, compress=lambda frames: frames, total=30, num_cond=2, num_pred=5,
                 lpips_fn=None):
    """city_sender.py:519-550 for one video.  x_gt (T,C,H,W) torch tensor in [0,1].  Returns (x_ge (total,C,H,W), d
    (total,), cycles): reconstructed frames, flags (1 = coded keyframe, 0 = predicted) and sampling cycles run."""
    total = min(total, x_gt.shape[0])
    gt = x_gt.unsqueeze(0)
    x_ge = compress(x_gt[:num_cond]).unsqueeze(0).clone()
    d = np.array([[1] * num_cond])
    cycles = 0
    while x_ge.shape[1] < total:
        l = x_ge.shape[1]
        d, x_ge = update(gt, x_ge, d, generate_frame, threshold, num_cond, num_pred, lpips_fn)
        cycles += 1
        if x_ge.shape[1] - l == 0:
            data_dec = compress(x_gt[l:l + num_cond]).unsqueeze(0)
            x_ge = torch.cat([x_ge, data_dec.to(x_ge.dtype)], dim=1)
            d = np.concatenate([d, np.array([[1] * data_dec.shape[1]])], axis=1)
    return x_ge[0, :total], d[0, :total], cycles
