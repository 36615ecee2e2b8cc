"""Import shim: the product package lives in the directory ``asr-study_b200/``
(the name the build contract asks for, which is not a valid Python identifier).
``import asr_study_b200.<x>`` resolves inside that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "asr-study_b200"))
from ._lib import lib, AsrError  # noqa: E402,F401
