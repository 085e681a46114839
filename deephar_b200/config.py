"""Mirror of deephar/config.py::ModelConfig (config.py:150-192), the pose layouts the hot
path needs (deephar/utils/pose.py:127-140: only num_joints / dim matter here) and the evaluation-time half of
DataConfig (config.py:6-50, 99-147) that the input pipeline (preprocess.py) reads."""


class DataConfig(object):
    """Input frame configuration (config.py:6-50).  Only the FIXED configuration -- what every evaluator uses
    (`dataconf.get_fixed_config()`, data/mpii.py:94-97) -- is functional here; the augmentation ranges are accepted and
    kept so that the reference's constructor calls work unchanged, but sampling from them is training-time code."""

    _FIXED = (('angle', 0), ('scale', 1), ('trans_x', 0), ('trans_y', 0), ('hflip', 0), ('chpower', 1),
              ('geoocclusion', None), ('subsampling', 1))

    def __init__(self, crop_resolution=(256, 256), image_channels=(3,), **kwargs):
        self.crop_resolution = tuple(crop_resolution)
        self.image_channels = tuple(image_channels)
        self.input_shape = self.crop_resolution + self.image_channels         # channels_last (config.py:3-4)
        for name, default in self._FIXED:
            setattr(self, 'fixed_' + name, kwargs.pop('fixed_' + name, default))
        self.augmentation = {k: kwargs.pop(k) for k in ('angles', 'scales', 'trans_x', 'trans_y', 'hflips', 'chpower',
                                                        'geoocclusion', 'subsampling') if k in kwargs}
        if kwargs:
            raise TypeError('DataConfig: unexpected argument(s) %s' % sorted(kwargs))

    def get_fixed_config(self):
        """config.py:42-50 (same keys)."""
        return {'angle': self.fixed_angle, 'scale': self.fixed_scale, 'transx': self.fixed_trans_x,
                'transy': self.fixed_trans_y, 'hflip': self.fixed_hflip, 'chpower': self.fixed_chpower,
                'geoocclusion': self.fixed_geoocclusion, 'subspl': self.fixed_subsampling}

    def random_data_generator(self):
        raise NotImplementedError('random augmentation (config.py:52-70) is training-time code: out of scope')


# the reference's instances (config.py:99-147, 195): evaluation only needs the resolution and the fixed sub-sampling
mpii_sp_dataconf = mpii_dataconf = DataConfig(crop_resolution=(256, 256))
pennaction_dataconf = DataConfig(crop_resolution=(256, 256), fixed_subsampling=6)
pennaction_pe_dataconf = DataConfig(crop_resolution=(256, 256))
human36m_dataconf = DataConfig(crop_resolution=(256, 256))
ntu_dataconf = DataConfig(crop_resolution=(256, 256), fixed_subsampling=4)
ntu_pe_dataconf = DataConfig(crop_resolution=(256, 256))


class ModelConfig(object):
    """Hyperparameters for models (same constructor as the reference)."""

    def __init__(self, input_shape, poselayout,
                 num_actions=[],
                 num_pyramids=8,
                 action_pyramids=[1, 2],
                 num_levels=4,
                 kernel_size=(5, 5),
                 growth=96,
                 image_div=8,
                 predict_rootz=False,
                 downsampling_type='maxpooling',
                 pose_replica=False,
                 num_pose_features=128,
                 num_visual_features=128,
                 sam_alpha=1,
                 dbg_decoupled_pose=False,
                 dbg_decoupled_h=False):
        self.input_shape = input_shape
        self.num_joints = poselayout.num_joints
        self.dim = poselayout.dim

        assert type(num_actions) == list, 'num_actions should be a list'
        self.num_actions = num_actions

        self.num_pyramids = num_pyramids
        self.action_pyramids = action_pyramids
        self.num_levels = num_levels
        self.kernel_size = kernel_size
        self.growth = growth
        self.image_div = image_div
        self.predict_rootz = predict_rootz
        self.downsampling_type = downsampling_type
        self.pose_replica = pose_replica
        self.num_pose_features = num_pose_features
        self.num_visual_features = num_visual_features
        self.sam_alpha = sam_alpha

        self.dbg_decoupled_pose = dbg_decoupled_pose
        self.dbg_decoupled_h = dbg_decoupled_h


class pa16j2d(object):
    num_joints = 16
    dim = 2


class pa16j3d(object):
    num_joints = 16
    dim = 3


class pa17j2d(object):
    num_joints = 17
    dim = 2


class pa17j3d(object):
    num_joints = 17
    dim = 3


class pa20j3d(object):
    num_joints = 20
    dim = 3


class pa21j3d(object):
    num_joints = 21
    dim = 3
