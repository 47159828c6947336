"""Device-side drop-ins for the reference's point-cloud export (create_pointcloud.py:65-105, utils/ply_utils.py:8-53).

`PLYSaver` keeps the reference's constructor, `add_depthmap(depth, image, intrinsics, extrinsics)` and `save(file)`; the
vertices stay in one growing device buffer (the reference does `.cpu().tolist()` per frame) and are written in the reference's
order.  `keep_mask` is the 33x33 dilation of the moving-object mask, `MaskVoter` the sliding-window vote of
create_pointcloud.py:80-104.  The arithmetic runs in libmonorec_b200.so (csrc/pointcloud.cu); no CPU fallback.
"""
import ctypes

import torch

from . import _lib


def keep_mask(cv_mask, mask_fill=32, thresh=0.1):
    """(conv2d(cv_mask >= thresh, ones(mask_fill+1), padding=mask_fill//2) < 1) as float  (create_pointcloud.py:77-78)."""
    if not cv_mask.is_cuda:
        raise _lib.MonorecLibraryError("monorec_b200.pointcloud needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    m = cv_mask.to(torch.float32).contiguous()
    B, _, H, W = m.shape
    out = torch.empty_like(m)
    with torch.cuda.device(m.device):
        _lib.check(lib.mr_pointcloud_keep_mask(m.data_ptr(), out.data_ptr(), B, H, W, int(mask_fill), float(thresh),
                                               torch.cuda.current_stream(m.device).cuda_stream), "mr_pointcloud_keep_mask")
    return out


class PLYSaver(torch.nn.Module):
    """Drop-in for utils/ply_utils.py:8-53."""

    def __init__(self, height, width, min_d=3, max_d=400, batch_size=1, roi=None, dropout=0):
        super().__init__()
        self.height, self.width = height, width
        self.min_d, self.max_d, self.roi, self.dropout = min_d, max_d, roi, dropout
        self._buf = None            # device float [capacity, 6]
        self._count = None          # device int64 [1]: vertices stored (negative: last add did not fit)

    def __len__(self):
        return 0 if self._count is None else int(self._count.item())

    @property
    def vertices(self):
        """Device tensor [N, 6] (x, y, z, red, green, blue)."""
        n = len(self)
        return self._buf[:n] if n else torch.empty(0, 6)

    def add_depthmap(self, depth, image, intrinsics, extrinsics, keep_masks=(), min_hits=1, rand=None):
        """depth: inverse depth [B,1,H,W] (the reference's argument name); keep_masks: the voting window's masks (optional:
        the reference multiplies the depth by the voted mask before calling; passing the masks here fuses that product)."""
        if not depth.is_cuda:
            raise _lib.MonorecLibraryError("monorec_b200.pointcloud needs CUDA tensors (no CPU fallback)")
        lib = _lib.load()
        dev = depth.device
        d = depth.to(torch.float32).contiguous()
        img = image.to(torch.float32).contiguous()
        K = intrinsics.to(torch.float32).contiguous()
        P = extrinsics.to(torch.float32).contiguous()
        B, _, H, W = d.shape
        masks = [m.to(torch.float32).contiguous() for m in keep_masks]
        if self.dropout > 0 and rand is None:
            rand = torch.rand_like(d)                                     # ply_utils.py:44-45
        if self._buf is None:
            self._buf = torch.empty(max(4 * B * H * W, 1 << 20), 6, device=dev)
            self._count = torch.zeros(1, dtype=torch.int64, device=dev)
        ws_bytes = lib.mr_pointcloud_workspace(B, H, W)
        ws = torch.empty(ws_bytes // 4, dtype=torch.int32, device=dev)
        roi = None if self.roi is None else (ctypes.c_int * 4)(*[int(v) for v in self.roi])
        # the buffer position of this batch is the count so far: one 8-byte D2H per batch (the reference copies every vertex)
        n_before = int(self._count.item())
        if n_before + B * H * W > self._buf.shape[0]:
            grown = torch.empty(2 * (n_before + B * H * W), 6, device=dev)
            grown[:n_before] = self._buf[:n_before]
            self._buf = grown
        with torch.cuda.device(dev):
            _lib.check(lib.mr_pointcloud_add(d.data_ptr(), img.data_ptr(), K.data_ptr(), P.data_ptr(),
                                             _lib.ptr_array(masks) if masks else None, len(masks), int(min_hits), B, H, W,
                                             float(self.min_d), float(self.max_d), roi,
                                             None if rand is None else rand.contiguous().data_ptr(), float(self.dropout),
                                             self._buf.data_ptr(), self._buf.shape[0], n_before, self._count.data_ptr(),
                                             ws.data_ptr(), ws_bytes, torch.cuda.current_stream(dev).cuda_stream),
                       "mr_pointcloud_add")

    def save(self, file):
        """Binary little-endian PLY, the reference's header (ply_utils.py:20-32)."""
        v = self.vertices.detach().to("cpu", torch.float32).contiguous()
        header = ("ply\nformat binary_little_endian 1.0\n"
                  f"element vertex {v.shape[0]}\n"
                  "property float x\nproperty float y\nproperty float z\n"
                  "property float red\nproperty float green\nproperty float blue\nend_header\n")
        file.write(header.encode(encoding="ascii"))
        file.write(v.numpy().tobytes())


class MaskVoter:
    """The sliding window of create_pointcloud.py:80-104: push one frame's tensors, get back the key frame (the middle of the
    window) with the window's keep masks once `buffer_length` frames are in."""

    def __init__(self, buffer_length=5, min_hits=1, mask_fill=32, thresh=0.1):
        self.buffer_length, self.min_hits, self.mask_fill, self.thresh = buffer_length, min_hits, mask_fill, thresh
        self.frames = []

    def push(self, result, data):
        out = result["result"]
        cvm = result["cv_mask"] if "cv_mask" in result else out.new_zeros(out.shape)
        self.frames.append((keep_mask(cvm, self.mask_fill, self.thresh), data["keyframe_pose"], data["keyframe_intrinsics"],
                            data["keyframe"], out))
        if len(self.frames) < self.buffer_length:
            return None
        key = self.frames[self.buffer_length // 2]
        masks = [f[0] for f in self.frames]
        del self.frames[0]
        return {"depth": key[4], "keyframe": key[3], "intrinsics": key[2], "pose": key[1], "keep_masks": masks,
                "min_hits": self.min_hits}
