"""Box / anchor / dataset utilities with the reference's names and signatures (reference utils.py), B200 side.

The numeric routines on the hot path (IoU, NMS, box refinement) run as CUDA kernels (cfun_b200.ops); the functions
here keep the reference call shapes: numpy in / numpy out for `non_max_suppression` and `compute_iou`, torch tensors
for `box_refinement` and `denorm_boxes_graph`.  Anchor generation is a one-off host computation.
"""
import numpy as np
import torch

from . import ops


# ---------------------------------------------------------------------------------------------------------
# boxes
# ---------------------------------------------------------------------------------------------------------
def extract_bboxes(mask):
    """mask [D,H,W,instances] -> int32 [instances,(z1,y1,x1,z2,y2,x2)] (reference utils.py:20-47: the same extent, taken
    over the whole mask, is reported for every instance; an extent that is flat in z yields zeros)."""
    n = mask.shape[-1]
    boxes = np.zeros((n, 6), dtype=np.int32)
    zy = np.where(np.sum(mask, axis=2) > 0)
    yx = np.where(np.sum(mask, axis=0) > 0)
    for i in range(n):
        z1, z2 = zy[0].min(), zy[0].max()
        y1, y2 = zy[1].min(), zy[1].max()
        x1, x2 = yx[1].min(), yx[1].max()
        boxes[i] = (z1, y1, x1, z2 + 1, y2 + 1, x2 + 1) if z1 != z2 else (0, 0, 0, 0, 0, 0)
    return boxes


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("cfun_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def compute_iou(box, boxes, box_volume, boxes_volume):
    """IoU of one box against many (reference utils.py:50-70); volumes are recomputed on device, the arguments are kept
    for signature compatibility."""
    b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float32)).to(_dev())
    a = torch.from_numpy(np.ascontiguousarray(box, dtype=np.float32).reshape(1, 6)).to(_dev())
    keep_eps = ops.iou_with_eps(a, b)
    return keep_eps[0].cpu().numpy()


def non_max_suppression(boxes, scores, threshold, max_num):
    """Greedy 3-D NMS (reference utils.py:122-157) on device: bitonic sort by (score desc, index asc), IoU bit matrix,
    sequential scan.  numpy in, numpy int32 out; accepts torch CUDA tensors too (then returns a CUDA int64 tensor)."""
    as_numpy = not torch.is_tensor(boxes)
    dev = _dev()
    b = torch.as_tensor(np.ascontiguousarray(boxes, dtype=np.float32) if as_numpy else boxes).to(dev).float()
    s = torch.as_tensor(np.ascontiguousarray(scores, dtype=np.float32) if as_numpy else scores).to(dev).float()
    if b.shape[0] == 0:
        return np.zeros(0, dtype=np.int32) if as_numpy else torch.zeros(0, dtype=torch.long, device=dev)
    order = ops.sort_desc(s)
    keep, count = ops.nms3d(b[order.long()], threshold, int(max_num))
    n = int(count.item())
    picked = order[keep[:n].long()]
    return picked.cpu().numpy().astype(np.int32) if as_numpy else picked.long()


def box_refinement(box, gt_box):
    """reference utils.py:92-119 (torch tensors in / out)."""
    return ops.box_refinement(box, gt_box)


def denorm_boxes_graph(boxes, size):
    d, h, w = size
    scale = torch.tensor([d, h, w, d, h, w], dtype=torch.float32, device=boxes.device)
    return boxes * scale


# ---------------------------------------------------------------------------------------------------------
# anchors (host, one-off; reference utils.py:467-528)
# ---------------------------------------------------------------------------------------------------------
def generate_anchors(scales, ratios, shape, feature_stride, anchor_stride):
    """Cube anchors centred on every `anchor_stride`-th cell of a [depth,height,width] feature map.

    Enumeration order is the reference's: np.meshgrid(z, y, x) with 'xy' indexing makes the flat order
    y-slowest, z-middle, x-fastest -- reproduced here explicitly rather than through meshgrid."""
    scales = np.atleast_1d(np.array(scales, dtype=np.float64))
    ratios = np.atleast_1d(np.array(ratios, dtype=np.float64))
    sizes = np.repeat(scales[None, :], len(ratios), axis=0).reshape(-1)      # every ratio keeps the cube (utils.py:486-489)
    cz = np.arange(0, shape[0], anchor_stride) * feature_stride
    cy = np.arange(0, shape[1], anchor_stride) * feature_stride
    cx = np.arange(0, shape[2], anchor_stride) * feature_stride
    centers = np.empty((len(cy), len(cz), len(cx), len(sizes), 3), dtype=np.float64)
    centers[..., 0] = cz[None, :, None, None]
    centers[..., 1] = cy[:, None, None, None]
    centers[..., 2] = cx[None, None, :, None]
    half = 0.5 * np.broadcast_to(sizes[None, None, None, :, None], centers.shape)
    return np.concatenate([centers - half, centers + half], axis=-1).reshape(-1, 6)


def generate_pyramid_anchors(scales, ratios, feature_shapes, feature_strides, anchor_stride):
    return np.concatenate([generate_anchors(scales[i], ratios, feature_shapes[i], feature_strides[i], anchor_stride)
                           for i in range(len(scales))], axis=0)


# ---------------------------------------------------------------------------------------------------------
# resizing (host side, off the hot path; scipy stands in for scikit-image, which the reference imports)
# ---------------------------------------------------------------------------------------------------------
def resize(image, output_shape, order=1, mode='constant', cval=0, clip=True, preserve_range=True, anti_aliasing=False,
           anti_aliasing_sigma=None):
    import scipy.ndimage as ndi
    image = np.asarray(image)
    out_shape = tuple(int(s) for s in output_shape)
    zoom = [o / i for o, i in zip(out_shape, image.shape)]
    return ndi.zoom(image.astype(np.float64), zoom, order=order, mode="grid-constant", cval=cval, grid_mode=True)


def resize_image(image, min_dim=None, max_dim=None, min_scale=None, mode="square"):
    """'none' and 'self' modes of reference utils.py:342-393 (the only ones its configs use)."""
    dtype = image.dtype
    h, w, d = image.shape[:3]
    window = (0, 0, 0, d, h, w)
    padding = [(0, 0)] * 4
    if mode == "none":
        return image, window, 1, padding, None
    if mode == "self":
        image = resize(image, (max_dim, max_dim, min_dim, 1), order=1, mode="constant", preserve_range=True)
        return image.astype(dtype), (0, 0, 0, min_dim, max_dim, max_dim), -1, padding, None
    raise NotImplementedError("IMAGE_RESIZE_MODE %r is not used by the CFUN configs" % mode)


def resize_mask(mask, scale, padding, max_dim=0, min_dim=0, crop=None, mode="square"):
    if mode == "self":
        mask = resize(mask, (max_dim, max_dim, min_dim), order=0, mode='constant', preserve_range=True)
        return np.round(mask).astype(np.int32)
    if mode == "none":
        return mask
    raise NotImplementedError(mode)


def unmold_mask(mask, bbox, image_shape):
    """reference utils.py:443-460: trilinear (align_corners=False) resize of the class-probability crop to the box
    size, pasted into a zero volume.  mask [d,h,w,classes] numpy, bbox (z1,y1,x1,z2,y2,x2)."""
    import torch.nn.functional as F
    z1, y1, x1, z2, y2, x2 = [int(v) for v in bbox]
    m = torch.from_numpy(np.ascontiguousarray(mask)).float().to(_dev()).permute(3, 0, 1, 2).unsqueeze(0)
    m = F.interpolate(m, size=(z2 - z1, y2 - y1, x2 - x1), mode='trilinear', align_corners=False)
    m = m.squeeze(0).permute(1, 2, 3, 0).cpu().numpy()
    full = np.zeros((image_shape[1], image_shape[2], image_shape[3], m.shape[-1]), dtype=np.float32)
    full[z1:z2, y1:y2, x1:x2, :] = m
    return full


# ---------------------------------------------------------------------------------------------------------
# dataset base class + metrics (surface that heart_main.py subclasses / calls; reference utils.py:181-315, 580-617)
# ---------------------------------------------------------------------------------------------------------
class Dataset(object):
    def __init__(self, class_map=None):
        self._image_ids = []
        self.image_info = []
        self.class_info = [{"source": "", "id": 0, "name": "BG"}]
        self.source_class_ids = {}

    def add_class(self, source, class_id, class_name):
        assert "." not in source, "Source name cannot contain a dot"
        if any(i["source"] == source and i["id"] == class_id for i in self.class_info):
            return
        self.class_info.append({"source": source, "id": class_id, "name": class_name})

    def add_image(self, source, image_id, path, **kwargs):
        info = {"id": image_id, "source": source, "path": path}
        info.update(kwargs)
        self.image_info.append(info)

    def image_reference(self, image_id):
        return ""

    def prepare(self, class_map=None):
        self.num_classes = len(self.class_info)
        self.class_ids = np.arange(self.num_classes)
        self.class_names = [c["name"].split(",")[0] for c in self.class_info]
        self.num_images = len(self.image_info)
        self._image_ids = np.arange(self.num_images)
        self.class_from_source_map = {"{}.{}".format(i['source'], i['id']): k for i, k in zip(self.class_info, self.class_ids)}
        self.sources = list(set(i['source'] for i in self.class_info))
        self.source_class_ids = {s: [k for k, i in enumerate(self.class_info) if k == 0 or i['source'] == s]
                                 for s in self.sources}

    def map_source_class_id(self, source_class_id):
        return self.class_from_source_map[source_class_id]

    def get_source_class_id(self, class_id, source):
        info = self.class_info[class_id]
        assert info['source'] == source
        return info['id']

    @property
    def image_ids(self):
        return self._image_ids

    def source_image_link(self, image_id):
        return self.image_info[image_id]["path"]

    def load_image(self, image_id):
        import nibabel as nib   # optional dependency of the data layer only
        image = nib.load(self.image_info[image_id]['path']).get_data().copy()
        return np.expand_dims(image, -1)

    def load_mask(self, image_id):
        return np.empty([0, 0, 0])


def compute_per_class_mask_iou(gt_masks, pred_masks):
    """[H,W,D,instances] masks -> per-instance IoU (reference utils.py:580-596)."""
    g = np.reshape(gt_masks > .5, (-1, gt_masks.shape[-1])).astype(np.float32)
    p = np.reshape(pred_masks > .5, (-1, pred_masks.shape[-1])).astype(np.float32)
    inter = np.sum(g * p, axis=0)
    union = g.sum(0) + p.sum(0) - inter
    return inter / (union + 1e-6)


def compute_mask_iou(gt_masks, pred_masks):
    g = (np.reshape(gt_masks, -1) > 0).astype(np.int64)
    p = (np.reshape(pred_masks, -1) > 0).astype(np.int64)
    inter = int(np.dot(g, p))
    return inter / (g.sum() + p.sum() - inter + 1e-6)
