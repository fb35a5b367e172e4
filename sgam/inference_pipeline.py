from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration  # noqa: F401
