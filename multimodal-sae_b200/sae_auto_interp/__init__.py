"""Drop-in mirror of the reference package name `sae_auto_interp` for the SAE hot path.

Only the modules on the encode / TopK / decode / activation-cache / top-activation-scan / steering / attribution path
exist here (`sae`, `features`, `config`, `utils`); they keep the reference's names, signatures and on-disk formats and
route the math to the sm_100a kernels of `saeb200`.  Normal use is the overlay `saeb200.dropin.install()`, which rebinds
an installed reference package's names to these classes; putting this directory's parent on PYTHONPATH ahead of the
reference also works for the cache / steering / attribution launchers (everything they import is here).
"""
