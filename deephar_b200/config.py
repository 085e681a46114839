"""Mirror of deephar/config.py::ModelConfig (config.py:150-192) and the pose layouts the hot
path needs (deephar/utils/pose.py:127-140: only num_joints / dim matter here)."""


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
