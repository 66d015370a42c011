"""API stub kept for `training_loop.py:107` (`grid_sample_gradfix.enabled = True`). grid_sample is only used by
the ADA augmentation pipe, which the LayoutDETR loss never applies (training/loss.py:68-73)."""
enabled = False
