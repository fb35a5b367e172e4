from sgam_neurips22_b200.warp import median_blur, render_projection_from_srcs_fast  # noqa: F401
