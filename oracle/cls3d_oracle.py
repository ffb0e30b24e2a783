"""CPU restatement (torch, float32 / autograd) of the 3-D neighbourhood regulariser — TEST INFRASTRUCTURE ONLY.

Follows loss_cls_3d, field_construction/utils/loss_utils.py:158-186, with the two random draws made explicit:
  down-sampling   :160-163   rows = randperm(N)[:max_points] when N > max_points            -> `downsample_idx`
  normalisation   :165-168   q = (p - min) / (max - min) over all elements, only if max > min
  samples         :171-173   randperm(n)[:sample_size]                                       -> `sample_idx`
  neighbours      :176-177   cdist + topk(k, largest=False); restated as EXACT squared distances, stable order (ties ->
                             lower index).  The reference's cdist takes the |x|^2 + |y|^2 - 2 x.y route, so on clouds whose
                             neighbour spacing approaches its rounding error its neighbour set is itself approximate; the
                             golden vectors (oracle/make_golden_cls3d.py) are checked to lie outside that regime.
  loss            :183-186   lambda * mean | q_s * (log(q_s + 1e-10) - log(q_nbr + 1e-10)) |
Never imported by langscene-x_b200/.
"""
import torch


def knn_exact(points, sample_idx, k):
    """(S, k) indices of the k nearest points of each sampled row (the row itself included), exact squared distances."""
    q = points[sample_idx]
    d2 = torch.zeros(q.shape[0], points.shape[0], dtype=points.dtype)
    for axis in range(3):                                     # coordinate by coordinate: no (S, N, 3) intermediate
        d2 += (q[:, axis, None] - points[None, :, axis]) ** 2
    return torch.sort(d2, dim=1, stable=True).indices[:, :k]


def loss_cls_3d(features, predictions, k, lambda_val, sample_idx, downsample_idx=None, neighbors=None):
    if downsample_idx is not None:
        features, predictions = features[downsample_idx], predictions[downsample_idx]
    lo, hi = predictions.min(), predictions.max()
    if bool(hi > lo):
        predictions = (predictions - lo) / (hi - lo)
    nbr = knn_exact(features.detach(), sample_idx, k) if neighbors is None else neighbors
    own = predictions[sample_idx].unsqueeze(1)
    other = predictions[nbr]
    kl = own * (torch.log(own + 1e-10) - torch.log(other + 1e-10))
    return lambda_val * kl.abs().mean(), nbr
