"""TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench's cpu_baseline): CPU restatement of the trainer's photometric loss
terms, main_avatar.py:193-222, for ONE view; a view batch is the mean over its views (all terms are means over equally
sized views).  The reference computes this inside AvatarTrainer.forward_one_pass, whose module cannot be imported here
(it pulls in packages this image lacks); the restatement below follows it line by line in torch and is PINNED by running that
method's unmodified source (oracle/ref_source.py compiles it out of the installed reference copy) on the same inputs:
tests/test_zz_loss_head.py::test_loss_oracle_pinned_by_reference_trainer_source (total, logged terms, gradients)."""
import torch


def photometric_terms(rgb_map, mask_map, color_img, mask_img, boundary_mask_img, bg_color):
    """rgb_map (H,W,3) rendered colour, mask_map (H,W,1) rendered opacity, color_img (H,W,3) ground truth,
    mask_img / boundary_mask_img (H,W) bool, bg_color (3,) -> (l1_loss, mask_loss) as the trainer computes them."""
    image = rgb_map.permute(2, 0, 1)                                               # :193
    color_img = color_img.clone()
    color_img[~mask_img] = bg_color                                                # :197
    gt_image = color_img.permute(2, 0, 1)                                          # :198
    mask = mask_img.to(rgb_map.dtype)                                              # :199
    boundary = 1.0 - boundary_mask_img.to(rgb_map.dtype)                           # :200
    image = image * boundary[None] + (1.0 - boundary[None]) * bg_color[:, None, None]        # :201
    gt_image = gt_image * boundary[None] + (1.0 - boundary[None]) * bg_color[:, None, None]  # :202
    l1_loss = torch.abs(image - gt_image).mean()                                   # :208
    rendered_mask = mask_map.squeeze(-1) * boundary                                # :215
    gt_mask = mask * boundary                                                      # :216
    mask_loss = torch.abs(rendered_mask - gt_mask).mean()                          # :220
    return l1_loss, mask_loss
