"""
ava_b200 -- B200-native implementation of the AVA VAE hot path (training/inference of
the 128x128-spectrogram VAE and the `get_spec` front end).  Mirrors the reference's
module layout for that path:

    ava.models.vae                 -> <this package>.models.vae
    ava.models.vae_dataset         -> <this package>.models.vae_dataset
    ava.models.window_vae_dataset  -> <this package>.models.window_vae_dataset
    ava.preprocessing.utils        -> <this package>.preprocessing.utils
    ava.preprocessing.preprocess   -> <this package>.preprocessing.preprocess  (process_sylls)
    ava.plotting.mmd_plots         -> <this package>.plotting.mmd_plots        (MMD^2 estimators)

The directory name contains hyphens, so import it with
``importlib.import_module("autoencoded-vocal-analysis_b200")`` or through the
``ava_b200`` alias module at the repository root.
"""
from ._lib import AvaB200Error, LIB_PATH, launch_count, lib  # noqa: F401

__version__ = "0.1.0"
