"""Install the drop-in modules under the reference's module name.

    import lama_b200.patch as p; p.install()      # BEFORE anything imports saicinpainting.training.modules
    # ... then run the reference unchanged, e.g. runpy.run_path("bin/predict.py")

``install()`` registers ``lama_b200.modules`` as ``sys.modules['saicinpainting.training.modules.ffc']``, so the
reference's own import sites bind to the replacement:

  * ``saicinpainting/training/modules/__init__.py:3``   ``from …ffc import FFCResNetGenerator``  (make_generator)
  * ``saicinpainting/training/modules/pix2pixhd.py:12`` ``from …ffc import FFCResnetBlock``
  * ``saicinpainting/evaluation/refinement.py:13``      ``from …ffc import FFCResnetBlock``      (isinstance check :271)

Ordering trap: the reference's ``ffc.py`` imports ``saicinpainting.training.modules.base``,
which executes the package ``__init__`` and binds the *reference* class by name.  ``lama_b200.modules`` therefore
never imports the reference package.  If the package was already imported, ``install()`` rebinds the three
import sites instead.  ``uninstall()`` restores everything (used by the tests).
"""
import importlib
import sys

_NAME = "saicinpainting.training.modules.ffc"
_SITES = [("saicinpainting.training.modules", ["FFCResNetGenerator"]),
          ("saicinpainting.training.modules.pix2pixhd", ["FFCResnetBlock"]),
          ("saicinpainting.evaluation.refinement", ["FFCResnetBlock"])]
_saved = {}


def install():
    from . import modules
    if _NAME in sys.modules and sys.modules[_NAME] is modules:
        return modules
    _saved.setdefault("module", sys.modules.get(_NAME))
    sys.modules[_NAME] = modules
    pkg = sys.modules.get("saicinpainting.training.modules")
    if pkg is not None:
        setattr(pkg, "ffc", modules)
    # the package (or a sibling) was imported before us: rebind the names it already copied
    for mod_name, names in _SITES:
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for n in names:
            if hasattr(mod, n) and getattr(mod, n) is not getattr(modules, n):
                _saved[(mod_name, n)] = getattr(mod, n)
                setattr(mod, n, getattr(modules, n))
    return modules


def uninstall():
    orig = _saved.pop("module", None)
    if orig is not None:
        sys.modules[_NAME] = orig
    else:
        sys.modules.pop(_NAME, None)
    for key in [k for k in _saved if isinstance(k, tuple)]:
        mod = sys.modules.get(key[0])
        if mod is not None:
            setattr(mod, key[1], _saved[key])
        _saved.pop(key)


def installed() -> bool:
    from . import modules
    return sys.modules.get(_NAME) is modules
