"""Drop-in replacements for the reference's `Matrix_Operators` and `Transforms` modules.

    import spectraldoublediffusiveconvection_b200.compat as compat
    compat.install()          # before `import Main`
    import Main               # the reference's drivers now run on the B200 path, unchanged

`install()` registers the two modules in sys.modules under the reference's bare module names, which is how
Main.py / Plot_Tools.py / Linear_Problem.py resolve them (function-local `from Matrix_Operators import ...`,
Main.py:97-98,193-194,230-232,453-456,760-763).
"""
import sys


def install():
    from . import Matrix_Operators, Transforms
    sys.modules["Matrix_Operators"] = Matrix_Operators
    sys.modules["Transforms"] = Transforms
    return Matrix_Operators, Transforms
