"""Drop-in replacements for pose_pipeline.wrappers.{mmpose,mmtrack,videopose3d} (same names, signatures, returns and
error behaviour as the reference functions)."""
