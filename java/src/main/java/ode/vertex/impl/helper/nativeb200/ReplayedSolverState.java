package ode.vertex.impl.helper.nativeb200;

import ode.solve.impl.util.SolverState;
import org.nd4j.linalg.api.ndarray.INDArray;
import org.nd4j.linalg.factory.Nd4j;

/**
 * The {@link SolverState} handed to {@link ode.solve.api.StepListener#step} when accepted steps are replayed from the
 * device-side trial log after a native solve.  The state vectors themselves stay on the device: listeners that only look
 * at time / step / error (StepCounter, the plotting listeners of the examples) work unchanged; the state accessors throw.
 * NOT COMPILED IN THIS REPOSITORY (no JDK / DL4J in the build image).
 */
final class ReplayedSolverState implements SolverState {

    private final float time;

    ReplayedSolverState(float time) {
        this.time = time;
    }

    @Override
    public INDArray getStateDot(long stage) {
        throw new UnsupportedOperationException("native ODE helper: stage derivatives stay on the device");
    }

    @Override
    public INDArray getCurrentState() {
        throw new UnsupportedOperationException("native ODE helper: the solver state stays on the device");
    }

    @Override
    public INDArray time() {
        return Nd4j.scalar(time);
    }

    @Override
    public double[] getInterpolationMidpoints() {
        // DormandPrince54Solver.java:44-47 (cMid); the device-side dense output uses the same coefficients
        return new double[]{6025192743d / 30085553152d / 2d, 0d, 51252292925d / 65400821598d / 2d, -2691868925d / 45128329728d / 2d,
                187940372067d / 1594534317056d / 2d, -1776094331d / 19743644256d / 2d, 11237099d / 235043384d / 2d};
    }
}
