"""Drop-in import surface: the module paths of the reference (`sgam.inference_pipeline`,
`sgam.generative_sensing_module.model`, `sgam.point_rendering.warp`, ...) re-exported from the B200 package
`sgam_neurips22_b200`, so the reference's main_scene_generation.py runs unchanged from this repository root."""
