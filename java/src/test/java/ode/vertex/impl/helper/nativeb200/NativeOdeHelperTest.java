package ode.vertex.impl.helper.nativeb200;

import ode.solve.conf.DormandPrince54Solver;
import ode.solve.conf.SolverConfig;
import ode.vertex.conf.OdeVertex;
import ode.vertex.conf.helper.FixedStep;
import ode.vertex.conf.helper.InputStep;
import org.deeplearning4j.nn.conf.ConvolutionMode;
import org.deeplearning4j.nn.conf.NeuralNetConfiguration;
import org.deeplearning4j.nn.conf.graph.GraphVertex;
import org.deeplearning4j.nn.conf.inputs.InputType;
import org.deeplearning4j.nn.conf.layers.BatchNormalization;
import org.deeplearning4j.nn.conf.layers.CnnLossLayer;
import org.deeplearning4j.nn.conf.layers.Convolution2D;
import org.deeplearning4j.nn.graph.ComputationGraph;
import org.junit.Assume;
import org.junit.Test;
import org.nd4j.linalg.activations.impl.ActivationIdentity;
import org.nd4j.linalg.api.ndarray.INDArray;
import org.nd4j.linalg.factory.Nd4j;
import org.nd4j.linalg.learning.config.Sgd;

import java.io.IOException;

import static org.junit.Assert.assertEquals;
import static org.junit.Assert.assertTrue;

/**
 * Reference-side tests of the native helper.  Mirrors src/test/java/ode/vertex/conf/OdeVertexTest.java:56-120 (the JSON shape
 * of an OdeVertex is unchanged by the native helper: conf classes keep fields and @JsonProperty names) and compares the
 * native solve with the reference's own Java path on the same graph, inputs and random-init weights (north_star: rtol 1e-4,
 * identical accepted / rejected counts).
 * <p>
 * NOT COMPILED OR RUN IN THIS REPOSITORY (no JDK / Maven / DL4J in the build image).  The numerical comparison it describes
 * is carried out against the C++ restatement of the reference (oracle/) by tests/test_gpu_parity.py through the same C ABI.
 */
public class NativeOdeHelperTest {

    private static GraphVertex vertex(ode.vertex.conf.helper.OdeHelper helper) {
        return new OdeVertex.Builder(new NeuralNetConfiguration.Builder(), "1", new BatchNormalization.Builder().nOut(3).build())
                .addLayer("2", new Convolution2D.Builder(3, 3).nOut(3).hasBias(false).activation(new ActivationIdentity())
                        .convolutionMode(ConvolutionMode.Same).build(), "1")
                .odeConf(helper)
                .build();
    }

    /** OdeVertexTest.serializeDeserialize with the patched conf classes: same JSON in, equal vertex out. */
    @Test
    public void serializeDeserialize() throws IOException {
        for (ode.vertex.conf.helper.OdeHelper helper : new ode.vertex.conf.helper.OdeHelper[]{
                new FixedStep(new DormandPrince54Solver(new SolverConfig(1e-3, 1e-3, 1e-10, 10)), Nd4j.arange(2)),
                new FixedStep(new DormandPrince54Solver(), Nd4j.linspace(0, 1, 5), true),
                new InputStep(new DormandPrince54Solver(), 1, true, true)}) {
            final GraphVertex v = vertex(helper);
            final String json = NeuralNetConfiguration.mapper().writeValueAsString(v);
            final OdeVertex back = NeuralNetConfiguration.mapper().readValue(json, OdeVertex.class);
            assertEquals("Not same!", v, back);
            assertEquals("Not same!", v.hashCode(), back.hashCode());
            assertEquals("Clone not equal!", v, v.clone());
            assertTrue("JSON must not mention the native helper", !json.contains("nativeb200"));
        }
    }

    /** Native forward + adjoint backward against the reference's Java path: states and gradients rtol 1e-4, equal nfe. */
    @Test
    public void nativeMatchesJavaPath() {
        Assume.assumeTrue("needs libnode4j_b200.so and a CUDA device", Node4j.deviceAvailable());
        final INDArray input = Nd4j.randn(new long[]{4, 3, 9, 9});
        final INDArray[] out = new INDArray[2];
        final INDArray[] grads = new INDArray[2];
        final int[] nfe = new int[2];
        for (int arm = 0; arm < 2; arm++) {
            System.setProperty("node4j.disable", arm == 0 ? "true" : "false");
            final ComputationGraph graph = new ComputationGraph(new NeuralNetConfiguration.Builder()
                    .updater(new Sgd()).seed(666).graphBuilder()
                    .setInputTypes(InputType.convolutional(9, 9, 3)).addInputs("input")
                    .addVertex("ode", vertex(new FixedStep(new DormandPrince54Solver(new SolverConfig(1e-3, 1e-3, 1e-20, 100)), Nd4j.arange(2))), "input")
                    .addLayer("output", new CnnLossLayer(), "ode").setOutputs("output").build());
            graph.init();
            out[arm] = graph.outputSingle(input).dup();
            graph.setInput(0, input);
            graph.setLabels(Nd4j.zeros(out[arm].shape()));
            graph.computeGradientAndScore();
            grads[arm] = graph.gradient().gradient().dup();
            nfe[arm] = ((ode.vertex.impl.OdeVertex) graph.getVertex("ode")).getFunction().getIterationCount();
        }
        assertTrue("final state", out[0].equalsWithEps(out[1], 1e-4 * out[0].amaxNumber().doubleValue()));
        assertTrue("gradients", grads[0].equalsWithEps(grads[1], 1e-4 * grads[0].amaxNumber().doubleValue()));
        assertEquals("nfe of the backward solve (3 + 6 * trials)", nfe[0], nfe[1]);
    }
}
