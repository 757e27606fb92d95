"""biogo_b200 -- B200-native replacement for the DP hot path of biogo's ``align`` package.

Only what the path needs lives here:

* ``csrc/``      hand-written sm_100a CUDA kernels + the C-ABI library (libbiogoalign.so)
* ``alphabet``   boundary types of biogo/alphabet used by align (Letters, QLetters, Alphabet)
* ``seq``        linear.Seq / linear.QSeq (the AlphabetSlicer arguments of Align)
* ``feat``       the feat.Pair result type
* ``matrix``     scoring tables of align/matrix used by the configs (Match, NUC_4, BLOSUM62)
* ``align``      SW / NW / Fitted / SWAffine / NWAffine / FittedAffine with the reference's
                 ``Align(reference, query)`` surface plus the new ``AlignBatch``

The host side is a thin mirror over the C ABI (ctypes); there is no CPU fallback: every
``Align`` call needs the CUDA library and a B200.
"""

__version__ = "0.1.0"
