"""Configuration attribute bag, name-compatible with reference config.py:15-232 (it is the reference's flag system:
class attributes overridden by sub-classes such as heart_main.HeartConfig, derived fields computed in __init__(stage))."""
import numpy as np


class Config(object):
    NAME = None
    GPU_COUNT = 1
    IMAGES_PER_GPU = 1
    STEPS_PER_EPOCH = 1000
    VALIDATION_STEPS = 50
    BACKBONE = "P3D131"
    BACKBONE_STRIDES = [4, 8, 16, 32, 64]
    BACKBONE_CHANNELS = [32, 64, 128, 256]
    FPN_CLASSIFY_FC_LAYERS_SIZE = 1024
    FPN_MASK_BRANCH_CHANNEL = 256
    TOP_DOWN_PYRAMID_SIZE = 256
    RPN_CONV_CHANNELS = 128
    NUM_CLASSES = 1
    RPN_ANCHOR_SCALES = (32, 64, 128, 256, 512)
    RPN_ANCHOR_RATIOS = [1]
    RPN_ANCHOR_STRIDE = 1
    RPN_NMS_THRESHOLD = 0.7
    RPN_TRAIN_ANCHORS_PER_IMAGE = 256
    POST_NMS_ROIS_TRAINING = 2000
    POST_NMS_ROIS_INFERENCE = 1000
    USE_MINI_MASK = True
    IMAGE_RESIZE_MODE = "square"
    IMAGE_MIN_DIM = 128
    IMAGE_MAX_DIM = 128
    IMAGE_MIN_SCALE = 0
    TRAIN_ROIS_PER_IMAGE = 200
    ROI_POSITIVE_RATIO = 0.33
    POOL_SIZE = [7, 7, 7]
    MASK_POOL_SIZE = [14, 14, 14]
    MAX_GT_INSTANCES = 100
    RPN_BBOX_STD_DEV = np.array([0.1, 0.1, 0.1, 0.2, 0.2, 0.2])
    BBOX_STD_DEV = np.array([0.1, 0.1, 0.1, 0.2, 0.2, 0.2])
    DETECTION_MAX_INSTANCES = 100
    DETECTION_MIN_CONFIDENCE = 0.7
    DETECTION_NMS_THRESHOLD = 0.3
    LEARNING_RATE = 0.001
    LEARNING_MOMENTUM = 0.9
    WEIGHT_DECAY = 0.0001
    LOSS_WEIGHTS = {"rpn_class_loss": 1., "rpn_bbox_loss": 1., "mrcnn_class_loss": 1., "mrcnn_bbox_loss": 1.,
                    "mrcnn_mask_loss": 1., "mrcnn_mask_edge_loss": 1.}
    USE_RPN_ROIS = True
    TRAIN_BN = False
    GRADIENT_CLIP_NORM = 5.0
    # switches between the two copies of the reference model (root = MM-WHS heart, LiTS_2017/ = liver); heart values here
    BACKBONE_STEM_KERNEL = (3, 7, 7)      # LiTS: (5, 7, 7)                    (LiTS_2017/backbone.py:124)
    UNET_DROPOUT = True                   # LiTS: no Dropout3d                 (LiTS_2017/mask_branch.py:19,130)
    MASK_CLASS_WEIGHT = None              # LiTS: CE weight [1, 1, 100]        (LiTS_2017/model.py:926)
    EDGE_LOSS_MODE = "magnitude"          # LiTS: "raw" Sobel responses        (LiTS_2017/model.py:967-975)
    STAGED_LOSSES = False                 # LiTS: detector losses in 'beginning', mask losses afterwards, detector frozen
    ROI_COUNT_ROUND = False               # LiTS: int(round(..)) RoI counts    (LiTS_2017/model.py:448,496)

    def __init__(self, stage):
        self.BATCH_SIZE = self.IMAGES_PER_GPU * self.GPU_COUNT
        lo, hi = self.IMAGE_MIN_DIM, self.IMAGE_MAX_DIM
        shape = {"crop": (lo, lo, lo), "self": (hi, hi, lo)}.get(self.IMAGE_RESIZE_MODE, (hi, hi, hi))
        self.IMAGE_SHAPE = np.array(list(shape) + [1])
        self.IMAGE_META_SIZE = 1 + 4 + 4 + 6 + 1 + self.NUM_CLASSES
        self.STAGE = stage
        side = 192 if stage == 'finetune' else 96
        self.MINI_MASK_SHAPE = (side,) * 3
        self.MASK_SHAPE = (side,) * 3
        self.DETECTION_TARGET_IOU_THRESHOLD = 0.5

    def display(self):
        print("\nConfigurations:")
        for a in dir(self):
            if not a.startswith("__") and not callable(getattr(self, a)):
                print("{:30} {}".format(a, getattr(self, a)))
        print("\n")


class HeartConfig(Config):
    """The MM-WHS configuration heart_main.py declares (reference heart_main.py:26-174), for callers that do not import
    heart_main itself (bench.py, tests)."""
    NAME = "heart"
    IMAGES_PER_GPU = 1
    NUM_CLASSES = 1 + 7
    STEPS_PER_EPOCH = 45
    VALIDATION_STEPS = 10
    BACKBONE = "P3D19"
    BACKBONE_STRIDES = [8, 16]
    BACKBONE_CHANNELS = [16, 32]
    FPN_CLASSIFY_FC_LAYERS_SIZE = 128
    UNET_MASK_BRANCH_CHANNEL = 20
    TOP_DOWN_PYRAMID_SIZE = 128
    RPN_CONV_CHANNELS = 256
    RPN_ANCHOR_SCALES = (64, 128)
    RPN_ANCHOR_STRIDE = 1
    RPN_ANCHOR_RATIOS = [1]
    RPN_TRAIN_ANCHORS_PER_IMAGE = 128
    PRE_NMS_LIMIT = 1000
    POST_NMS_ROIS_TRAINING = 500
    POST_NMS_ROIS_INFERENCE = 64
    USE_MINI_MASK = False
    IMAGE_RESIZE_MODE = "self"
    IMAGE_MIN_DIM = 192
    IMAGE_MAX_DIM = 320
    IMAGE_MIN_SCALE = 0
    IMAGE_CHANNEL_COUNT = 1
    TRAIN_ROIS_PER_IMAGE = 15
    POOL_SIZE = [12, 12, 12]
    MASK_POOL_SIZE = [96, 96, 96]
    DETECTION_MIN_CONFIDENCE = 0.7
    DETECTION_NMS_THRESHOLD = 0.3
    MAX_GT_INSTANCES = 32
    DETECTION_MAX_INSTANCES = 32
    LOSS_WEIGHTS = {"rpn_class_loss": 100., "rpn_bbox_loss": 50., "mrcnn_class_loss": 1., "mrcnn_bbox_loss": 20.,
                    "mrcnn_mask_loss": 1., "mrcnn_mask_edge_loss": 1.}
    TRAIN_BN = False


class LiTSConfig(Config):
    """The liver configuration LiTS_main.py declares (reference LiTS_2017/LiTS_main.py:28-121) with the stage-dependent
    fields of LiTS_2017/config.py:197-226: BASELINE.json config 3."""
    NAME = "LiTS"
    IMAGES_PER_GPU = 1
    NUM_CLASSES = 1 + 2
    STEPS_PER_EPOCH = 100
    VALIDATION_STEPS = 20
    BACKBONE = "P3D35"
    BACKBONE_STRIDES = [8, 16]
    BACKBONE_CHANNELS = [24, 48]
    BACKBONE_STEM_KERNEL = (5, 7, 7)
    FPN_CLASSIFY_FC_LAYERS_SIZE = 320
    UNET_MASK_BRANCH_CHANNEL = 32
    TOP_DOWN_PYRAMID_SIZE = 160
    RPN_CONV_CHANNELS = 320
    RPN_ANCHOR_SCALES = (64, 128)
    RPN_ANCHOR_STRIDE = 1
    RPN_ANCHOR_RATIOS = [1]
    RPN_TRAIN_ANCHORS_PER_IMAGE = 128
    PRE_NMS_LIMIT = 1000
    POST_NMS_ROIS_TRAINING = 500
    POST_NMS_ROIS_INFERENCE = 50
    USE_MINI_MASK = False
    IMAGE_RESIZE_MODE = "self"
    IMAGE_MIN_DIM = 256
    IMAGE_MAX_DIM = 320
    IMAGE_MIN_SCALE = 0
    IMAGE_CHANNEL_COUNT = 1
    POOL_SIZE = [12, 12, 12]
    MASK_POOL_SIZE = [32, 80, 80]
    DETECTION_MIN_CONFIDENCE = 0.7
    DETECTION_NMS_THRESHOLD = 0.7
    MAX_GT_INSTANCES = 32
    DETECTION_MAX_INSTANCES = 32
    LOSS_WEIGHTS = {"rpn_class_loss": 50., "rpn_bbox_loss": 5., "mrcnn_class_loss": 50., "mrcnn_bbox_loss": 5.,
                    "mrcnn_mask_loss": 2., "mrcnn_mask_edge_loss": 0.25}
    UNET_DROPOUT = False
    MASK_CLASS_WEIGHT = [1., 1., 100.]
    EDGE_LOSS_MODE = "raw"
    STAGED_LOSSES = True
    ROI_COUNT_ROUND = True

    def __init__(self, stage):
        super().__init__(stage)
        self.IMAGE_SHAPE = np.array([self.IMAGE_MAX_DIM, self.IMAGE_MAX_DIM, self.IMAGE_MIN_DIM, 1])
        mp = tuple(int(v) for v in self.MASK_POOL_SIZE)
        self.MASK_SHAPE = tuple(2 * v for v in mp) if stage == 'finetune' else mp
        self.MINI_MASK_SHAPE = self.MASK_SHAPE
        if stage == 'beginning':
            self.TRAIN_ROIS_PER_IMAGE, self.ROI_POSITIVE_RATIO = 50, 0.33
        else:                                    # 'together' / 'finetune': positives only
            self.TRAIN_ROIS_PER_IMAGE, self.ROI_POSITIVE_RATIO = 4, 1.


def lits_config(stage="finetune", image_min=256, image_max=320, mask_pool=(32, 80, 80), **overrides):
    """LiTSConfig at a given input size / mask crop (BASELINE config 3 uses the defaults)."""
    attrs = dict(IMAGE_MIN_DIM=image_min, IMAGE_MAX_DIM=image_max, MASK_POOL_SIZE=[int(v) for v in mask_pool])
    attrs.update(overrides)
    return type("LiTSConfig%dx%d" % (image_min, image_max), (LiTSConfig,), attrs)(stage)


def heart_config(image_dim=256, stage="beginning", mask_pool=96, anchor_scales=(64, 128), **overrides):
    """HeartConfig at a cubic input size (BASELINE.json config 2 uses 256), optionally with a smaller mask crop."""
    attrs = dict(IMAGE_MIN_DIM=image_dim, IMAGE_MAX_DIM=image_dim, MASK_POOL_SIZE=[mask_pool] * 3,
                 RPN_ANCHOR_SCALES=tuple(anchor_scales))
    attrs.update(overrides)
    cls = type("HeartConfig%d" % image_dim, (HeartConfig,), attrs)
    cfg = cls(stage)
    side = mask_pool * (2 if stage == "finetune" else 1)
    cfg.MASK_SHAPE = (side,) * 3
    cfg.MINI_MASK_SHAPE = cfg.MASK_SHAPE
    return cfg
