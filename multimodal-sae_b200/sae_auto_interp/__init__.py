"""Drop-in mirror of the reference package name `sae_auto_interp` for the SAE hot path.

Only the modules on the encode / TopK / decode / activation-cache / top-activation-scan / steering path exist here
(`sae`, `features`, `config`, `utils`); they keep the reference's names, signatures and on-disk formats and route the
math to the sm_100a kernels of `saeb200`.  Put this directory's parent on PYTHONPATH ahead of the reference to switch
the existing `cache_image` / `explain` / `steering` launchers over.
"""
