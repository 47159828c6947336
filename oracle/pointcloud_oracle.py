"""CPU restatement (torch) of the reference's point-cloud export -- TEST INFRASTRUCTURE, never imported by the product.

Follows create_pointcloud.py:77-78 (dilated keep mask), :93-95 (window vote, depth *= mask) and utils/ply_utils.py:34-53
(PLYSaver.add_depthmap with model/layers.py:43-58 Backprojection).  Pinned on tests/golden/pointcloud.npz, which
tests/golden/make_golden.py --only-pointcloud writes with the unmodified PLYSaver.
"""
import torch
import torch.nn.functional as F


def keep_mask(cv_mask, mask_fill=32, thresh=0.1):
    mask = (cv_mask >= thresh).to(torch.float32)                                                  # :77
    return (F.conv2d(mask, mask.new_ones((1, 1, mask_fill + 1, mask_fill + 1)), padding=mask_fill // 2) < 1).to(torch.float32)


def add_depthmap(inv_depth, image, intrinsics, pose, keep_masks=(), min_hits=1, min_d=3, max_d=400, roi=None, dropout=0.0,
                 rand=None):
    """-> [N, 6] vertices in the reference's order."""
    depth = inv_depth.clone()
    if len(keep_masks):                                                                           # :93-95
        voted = (torch.sum(torch.stack(list(keep_masks)), dim=0) > len(keep_masks) - min_hits).to(torch.float32)
        depth = depth * voted
    depth = 1 / depth                                                                             # ply_utils.py:36
    img = (image + .5) * 255
    mask = (min_d <= depth) & (depth <= max_d)
    if roi is not None:
        mask[:, :, :roi[0], :] = False
        mask[:, :, roi[1]:, :] = False
        mask[:, :, :, :roi[2]] = False
        mask[:, :, :, roi[3]:] = False
    if dropout > 0:
        mask = mask & (rand > dropout)
    B, _, H, W = depth.shape
    yy, xx = torch.meshgrid([torch.arange(0., float(H)), torch.arange(0., float(W))], indexing="ij")
    coord = torch.stack([xx.reshape(-1), yy.reshape(-1), torch.ones(H * W)], 0).unsqueeze(0).repeat(B, 1, 1)
    cam = torch.matmul(torch.inverse(intrinsics)[:, :3, :3], coord) * depth.view(B, 1, -1)        # layers.py:56-57
    cam = torch.cat([cam, torch.ones(B, 1, H * W)], 1)
    world = (pose @ cam)[:, :3, :]
    data = torch.cat([world, img.view_as(world)], dim=1).permute(0, 2, 1)
    return data[mask.view(B, 1, -1).permute(0, 2, 1).expand(-1, -1, 6)].view(-1, 6)
