"""neuralode4j_b200 — B200-native ODE-block engine behind neuralODE4j's OdeVertex API.

Only the hot path of the reference lives here: DormandPrince54 adaptive integration of an OdeVertex inner graph
(forward) and the adjoint backward solve, as hand-written sm_100a CUDA behind a plain C ABI (include/node4j.h).
This package is the host-side mirror of the reference's configuration/vertex interface for that path.
"""
from .conf import (ActivationLayer, BatchNormalization, Convolution2D, DenseLayer, DormandPrince54Solver, FixedStep,
                   InputStep, InputType, Layer, Mask, MergeVertex, ShapeMatchVertex, SolverConfig, StepCounter, StepListener)
from .vertex import OdeVertex, OdeVertexImpl
from . import _lib

__all__ = ["ActivationLayer", "BatchNormalization", "Convolution2D", "DenseLayer", "DormandPrince54Solver", "FixedStep",
           "InputStep", "InputType", "Layer", "Mask", "MergeVertex", "ShapeMatchVertex", "SolverConfig", "StepCounter", "StepListener", "OdeVertex", "OdeVertexImpl"]
