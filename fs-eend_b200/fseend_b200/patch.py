"""Opt-in hook that puts the device helpers underneath the REFERENCE's own modules.

The reference imports ``datasets.feature``, ``train.utils.loss`` and ``train.utils.make_rttm`` from namespace directories
(no ``__init__.py``), so this package must not ship directories with those names: they would shadow the reference's
``datasets.feature.extract_fbank`` / ``train.oln_tfm_enc_dec`` / ``datasets.diarization_dataset`` regardless of the
``sys.path`` order (round-1 review).  The helpers therefore live here (``fseend_b200.feature/.loss/.rttm``) and a user
who wants them under the reference's names calls ``patch_reference()`` AFTER the reference modules are importable:

    import fseend_b200.patch as P; P.patch_reference()           # replaces three functions, nothing else

``train.utils.loss.standard_loss`` (reference train/utils/loss.py:119-125), the PIT losses (``batch_pit_loss`` :98-116,
``batch_pit_n_speaker_loss`` :257-327, ``..._label_delay`` :329-403), ``train.utils.make_rttm.make_rttm`` (:10-28) are
replaced; ``datasets.feature`` gains ``splice_subsample`` (splice :111-133 + subsample :103-108 fused) and
keeps every reference function (``extract_fbank``, ``stft`` ... are untouched).
"""
import importlib


def patch_reference(loss: bool = True, rttm: bool = True, feature: bool = True):
    """Returns the list of ``module.attribute`` names that were replaced (modules that cannot be imported are skipped)."""
    done = []

    def _try(modname):
        try:
            return importlib.import_module(modname)
        except Exception:
            return None

    if loss:
        m = _try("train.utils.loss")
        if m is not None:
            from . import loss as L
            for fn in ("standard_loss", "prepare_labels", "batch_pit_loss", "batch_pit_n_speaker_loss",
                       "batch_pit_n_speaker_loss_label_delay"):
                setattr(m, fn, getattr(L, fn))
                done.append("train.utils.loss." + fn)
    if rttm:
        m = _try("train.utils.make_rttm")
        if m is not None:
            from . import rttm as R
            m.make_rttm = R.make_rttm
            done.append("train.utils.make_rttm.make_rttm")
    if feature:
        m = _try("datasets.feature")
        if m is not None:
            from . import feature as F
            m.splice_subsample = F.splice_subsample
            done.append("datasets.feature.splice_subsample")
    return done
