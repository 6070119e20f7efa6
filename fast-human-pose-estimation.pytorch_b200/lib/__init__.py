"""Mirror of the reference's `lib/` package for the hot path only: models.hourglass / models.pose_hrnet
(get_pose_net + forward), core.loss (JointsMSELoss), core.function (train / fpd_train / validate),
core.inference (get_max_preds), utils.transforms (flip_back), nms (gpu_nms). Put this directory ahead of the
reference's `lib` on sys.path (tools/_init_paths.py) and tools/fpd_train.py resolves to these."""
