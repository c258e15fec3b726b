"""Enumerations of the reference plugin API (`HE Wrapper/IVector.cs:15-18`, `IMatrix.cs:14-17`)."""
import enum


class EVectorFormat(enum.IntEnum):
    dense = 0
    sparse = 1


class EMatrixFormat(enum.IntEnum):
    ColumnMajor = 0
    RowMajor = 1
