"""Action-recognition heads on top of the drop-in encoder.

Mirror of `lib/model/model_action.py` -- `ActionHeadClassification` (:6-29), `ActionHeadEmbed` (:31-48), `ActionNet`
(:50-71): same constructors, same sub-module names, so `state_dict`s are interchangeable with the reference's checkpoints
(`train_action.py:213-219` key 'model').  Both heads read the representation ONLY through its mean over the T frames
(:20-21 / :43-44), so whenever nothing has to be differentiated and the head's dropout is inactive, `ActionNet.forward`
asks the encoder for `get_representation_pooled`: the tail GEMM's epilogue accumulates the temporal mean and the
(N*M, T, J, 512) representation -- 4.3 GB for a train_action.py batch of 128 two-person clips -- is never written.
Otherwise the reference's op sequence runs unchanged on `get_representation`.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class ActionHeadClassification(nn.Module):
    def __init__(self, dropout_ratio=0., dim_rep=512, num_classes=60, num_joints=17, hidden_dim=2048):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout_ratio)
        self.bn = nn.BatchNorm1d(hidden_dim, momentum=0.1)
        self.relu = nn.ReLU(inplace=True)
        self.fc1 = nn.Linear(dim_rep * num_joints, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, num_classes)

    def forward_pooled(self, pooled):
        """pooled: (N, M, J, C) = mean over T of the representation."""
        N, M = pooled.shape[:2]
        feat = pooled.reshape(N, M, -1).mean(dim=1)          # model_action.py:22-23
        return self.fc2(self.relu(self.bn(self.fc1(feat))))  # :24-27

    def forward(self, feat):
        """feat: (N, M, T, J, C)  (model_action.py:15-28)"""
        feat = self.dropout(feat)
        return self.forward_pooled(feat.permute(0, 1, 3, 4, 2).mean(dim=-1))


class ActionHeadEmbed(nn.Module):
    def __init__(self, dropout_ratio=0., dim_rep=512, num_joints=17, hidden_dim=2048):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout_ratio)
        self.fc1 = nn.Linear(dim_rep * num_joints, hidden_dim)

    def forward_pooled(self, pooled):
        N, M = pooled.shape[:2]
        feat = self.fc1(pooled.reshape(N, M, -1).mean(dim=1))
        return F.normalize(feat, dim=-1)                       # model_action.py:46-47

    def forward(self, feat):
        feat = self.dropout(feat)
        return self.forward_pooled(feat.permute(0, 1, 3, 4, 2).mean(dim=-1))


class ActionNet(nn.Module):
    def __init__(self, backbone, dim_rep=512, num_classes=60, dropout_ratio=0., version='class', hidden_dim=2048,
                 num_joints=17):
        super().__init__()
        self.backbone = backbone
        self.feat_J = num_joints
        if version == 'class':
            self.head = ActionHeadClassification(dropout_ratio=dropout_ratio, dim_rep=dim_rep, num_classes=num_classes,
                                                 num_joints=num_joints)
        elif version == 'embed':
            self.head = ActionHeadEmbed(dropout_ratio=dropout_ratio, dim_rep=dim_rep, hidden_dim=hidden_dim,
                                        num_joints=num_joints)
        else:
            raise Exception('Version Error.')

    def _can_pool(self, x):
        drop_active = self.training and self.head.dropout.p > 0
        # (nn.DataParallel replicas keep their parameters as plain attributes: ask the encoder, not .parameters())
        ps = self.backbone._ordered_params() if hasattr(self.backbone, "_ordered_params") else list(self.backbone.parameters())
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p is not None and p.requires_grad for p in ps))
        return (not drop_active) and (not needs_grad) and hasattr(self.backbone, "get_representation_pooled") \
            and not (self.backbone.training and self.backbone._drop_path_scale(1, 1, x.device) is not None)

    def forward(self, x):
        """x: (N, M, T, 17, 3)  (model_action.py:62-71)"""
        N, M, T, J, C = x.shape
        x = x.reshape(N * M, T, J, C)
        if self._can_pool(x):
            pooled = self.backbone.get_representation_pooled(x)              # (N*M, J, dim_rep)
            return self.head.forward_pooled(pooled.reshape(N, M, self.feat_J, -1))
        feat = self.backbone.get_representation(x)
        feat = feat.reshape([N, M, T, self.feat_J, -1])
        return self.head(feat)
