"""ScoreNetwork -- drop-in for multi_model/score_network.py:9-53 (same constructor, forward signature, outputs
and state-dict keys `extrat_featurePN2.*`)."""
import torch.nn as nn

from .pointnet2 import PointNet2Seg


class ScoreNetwork(nn.Module):
    def __init__(self, training=True, k_obj=2):
        super().__init__()
        self.is_training = training
        self.k_obj = k_obj
        self.extrat_featurePN2 = PointNet2Seg(input_chann=6, k_score=1, k_obj=self.k_obj)
        self.criterion_cls = nn.NLLLoss(reduction='mean')
        self.criterion_reg = nn.MSELoss(reduction='mean')

    def compute_loss(self, pscore, tscore):
        """MSE between predicted and target per-point grasp score, both (B,N)."""
        return self.criterion_reg(pscore, tscore.float())

    def prefetch(self, pc):
        """Optional (not in the reference): announce the NEXT batch so its FPS / ball-query / 3-NN chain overlaps the
        MLPs of the batch whose forward() is called next.  `pc` must be a contiguous (B,N,6) float32 CUDA tensor and
        the same tensor must be passed to forward() afterwards."""
        self.extrat_featurePN2.prefetch(pc)

    def join_prefetch(self):
        """Optional: make the current stream wait for every outstanding prefetch()."""
        self.extrat_featurePN2.join_prefetch()

    def forward(self, pc, pc_score=None, pc_label=None):
        """pc (B,N,>=6) [, pc_score (B,N)] -> (all_feature (B,N,256), output_score (B,N), loss | None).
        all_feature is the 256-channel output of the last feature-propagation layer (pointnet2.py:121), not the
        128-channel head feature the reference's docstring mentions."""
        feature, output_score = self.extrat_featurePN2(pc[:, :, :6].permute(0, 2, 1))
        all_feature = feature.transpose(2, 1)
        loss = None
        if self.is_training and pc_score is not None:
            loss = self.compute_loss(output_score, pc_score)
        return all_feature, output_score, loss
