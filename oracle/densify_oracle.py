"""TEST INFRASTRUCTURE — CPU restatement (torch, fp32) of the reference's adaptive density control.  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs may import this; the product never does.

Restates, on the reference's own tensor shapes (six parameter groups + their Adam moments):
  densify_and_prune        LG/scene/gaussian_model.py:456-476 (clone :437-454, split :411-435, prune_points :347-363,
                           _prune_optimizer :331-345, cat_tensors_to_optimizer :365-385, densification_postfix :387-409)
  add_densification_stats  :478-480 with the max_radii2D update of LG/train.py:268
  reset_opacity            :258-261 (replace_tensor_to_optimizer :316-329)
  activations              :36-50,102-130 and their gradients (what autograd does for the reference)
as ONE index plan (which source row every output row comes from) instead of the reference's four rounds of boolean
indexing, so that it is an independent statement of the same result rather than a transcription.

PINNED against the reference class itself, executed in the build container (tests/golden/make_densify_golden.py ->
tests/golden/densify_reference.npz): identical row counts and order, parameters / moments to 1e-6.
"""
import torch

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def build_rotation(r):
    """utils/general_utils.py:78-99"""
    n = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / n[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, 1).view(-1, 3, 3)


def plan(params, xyz_gradient_accum, denom, max_grad, min_opacity, extent, max_screen_size, percent_dense):
    """-> (src rows, kind per output row [0 original, 1 clone, 2 first child, 3 second child], split parents)"""
    P = params["xyz"].shape[0]
    grads = (xyz_gradient_accum / denom).reshape(P)
    grads = torch.where(torch.isnan(grads), torch.zeros_like(grads), grads)          # :457-458
    scale = torch.exp(params["scaling"])
    smax = scale.max(dim=1).values
    hot = grads.abs() >= max_grad                                                    # :416 / :439
    dense = percent_dense * extent
    clone = hot & (smax <= dense)                                                    # :440-441
    split = hot & (smax > dense)                                                     # :417-418
    faint = torch.sigmoid(params["opacity"]).reshape(P) < min_opacity                # :464
    child_scale = torch.exp(torch.log(scale / (0.8 * 2)))                            # :425 then get_scaling
    if max_screen_size:                                                              # :465-468 (max_radii2D is all
        big_self = smax > 0.1 * extent                                               # zeros by then, :409: never true)
        big_child = child_scale.max(dim=1).values > 0.1 * extent
    else:
        big_self = big_child = torch.zeros(P, dtype=torch.bool)
    idx = torch.arange(P)
    keep = idx[~split & ~(faint | big_self)]
    cloned = idx[clone & ~(faint | big_self)]
    parents = idx[split]
    child_ok = ~(faint | big_child)[parents]
    kids = parents[child_ok]
    src = torch.cat([keep, cloned, kids, kids])
    kind = torch.cat([torch.zeros_like(keep), torch.ones_like(cloned), torch.full_like(kids, 2),
                      torch.full_like(kids, 3)])
    # sample row of every child: the j-th split parent draws rows j and S + j (:421-423, repeat(N,1))
    j = torch.arange(parents.numel())[child_ok]
    eps_row = torch.cat([torch.zeros_like(keep), torch.zeros_like(cloned), j, parents.numel() + j])
    return src, kind, eps_row, parents


def densify_and_prune(params, exp_avg, exp_avg_sq, xyz_gradient_accum, denom, eps, max_grad, min_opacity, extent,
                      max_screen_size, percent_dense=0.01):
    """params / exp_avg / exp_avg_sq: dicts over GROUPS in the reference's shapes; eps (2S,3) unit normals.
    Returns the new (params, exp_avg, exp_avg_sq) and the counts."""
    src, kind, eps_row, parents = plan(params, xyz_gradient_accum, denom, max_grad, min_opacity, extent,
                                       max_screen_size, percent_dense)
    assert eps.shape[0] == 2 * parents.numel()
    child = kind >= 2
    out_p, out_m, out_v = {}, {}, {}
    for k in GROUPS:
        out_p[k] = params[k][src].clone()
        fresh = (kind != 0).view(-1, *([1] * (params[k].dim() - 1)))
        out_m[k] = torch.where(fresh, torch.zeros((), dtype=torch.float32), exp_avg[k][src])   # :376-378
        out_v[k] = torch.where(fresh, torch.zeros((), dtype=torch.float32), exp_avg_sq[k][src])
    s = torch.exp(params["scaling"][src[child]])
    samples = eps[eps_row[child]] * s                                                # torch.normal(0, std), :421-423
    R = build_rotation(params["rotation"][src[child]])
    out_p["xyz"][child] = torch.bmm(R, samples.unsqueeze(-1)).squeeze(-1) + params["xyz"][src[child]]   # :424
    out_p["scaling"][child] = torch.log(s / (0.8 * 2))                               # :425
    counts = dict(kept=int((kind == 0).sum()), cloned=int((kind == 1).sum()), split_parents=int(parents.numel()),
                  split_kept=int((kind == 2).sum()), P=int(src.numel()))
    return out_p, out_m, out_v, counts


def add_densification_stats(xyz_gradient_accum, denom, max_radii2D, grad2D, radii):
    vis = radii > 0
    accum, den, mr = xyz_gradient_accum.clone().reshape(-1), denom.clone().reshape(-1), max_radii2D.clone()
    mr[vis] = torch.max(mr[vis], radii[vis].float())                                 # LG/train.py:268
    accum[vis] += torch.norm(grad2D[vis, :2], dim=-1)                                # :479
    den[vis] += 1                                                                    # :480
    return accum, den, mr


def reset_opacity(opacity):
    o = torch.min(torch.sigmoid(opacity), torch.ones_like(opacity) * 0.01)           # :259
    return torch.log(o / (1 - o))


def activations(opacity, scaling, rotation):
    """get_opacity / get_scaling / get_rotation (:102-130)"""
    return torch.sigmoid(opacity), torch.exp(scaling), torch.nn.functional.normalize(rotation)
