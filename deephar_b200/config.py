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
    """Hyperparameters of the SPNet models: the reference's constructor (config.py:150-192 -- names, order and defaults
    are the interface every experiment script uses), every argument kept as an attribute of the same name."""

    _STORED = ('input_shape', 'num_actions', 'num_pyramids', 'action_pyramids', 'num_levels', 'kernel_size', 'growth',
               'image_div', 'predict_rootz', 'downsampling_type', 'pose_replica', 'num_pose_features',
               'num_visual_features', 'sam_alpha', 'dbg_decoupled_pose', 'dbg_decoupled_h')

    def __init__(self, input_shape, poselayout, num_actions=[], num_pyramids=8, action_pyramids=[1, 2], num_levels=4,
                 kernel_size=(5, 5), growth=96, image_div=8, predict_rootz=False, downsampling_type='maxpooling',
                 pose_replica=False, num_pose_features=128, num_visual_features=128, sam_alpha=1,
                 dbg_decoupled_pose=False, dbg_decoupled_h=False):
        given = locals()
        if type(num_actions) != list:
            raise AssertionError('num_actions should be a list')
        for name in self._STORED:
            setattr(self, name, given[name])
        self.num_joints, self.dim = poselayout.num_joints, poselayout.dim          # all the forward path reads of a layout


def _layout(name, num_joints, dim):
    """deephar/utils/pose.py:127-140: of a pose layout the models only read the joint count and the dimension."""
    return type(name, (object,), {'num_joints': num_joints, 'dim': dim, '__doc__': '%d joints, %d-D' % (num_joints, dim)})


pa16j2d, pa16j3d = _layout('pa16j2d', 16, 2), _layout('pa16j3d', 16, 3)
pa17j2d, pa17j3d = _layout('pa17j2d', 17, 2), _layout('pa17j3d', 17, 3)
pa20j3d, pa21j3d = _layout('pa20j3d', 20, 3), _layout('pa21j3d', 21, 3)
