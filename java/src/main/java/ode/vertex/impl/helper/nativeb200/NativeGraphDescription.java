package ode.vertex.impl.helper.nativeb200;

import ode.vertex.impl.ShapeMatchVertex;
import org.deeplearning4j.nn.api.Layer;
import org.deeplearning4j.nn.conf.ConvolutionMode;
import org.deeplearning4j.nn.conf.layers.ActivationLayer;
import org.deeplearning4j.nn.conf.layers.BatchNormalization;
import org.deeplearning4j.nn.conf.layers.ConvolutionLayer;
import org.deeplearning4j.nn.conf.layers.DenseLayer;
import org.deeplearning4j.nn.graph.ComputationGraph;
import org.deeplearning4j.nn.graph.vertex.GraphVertex;
import org.deeplearning4j.nn.graph.vertex.impl.LayerVertex;
import org.nd4j.linalg.activations.IActivation;
import org.nd4j.linalg.activations.impl.ActivationELU;
import org.nd4j.linalg.activations.impl.ActivationIdentity;
import org.nd4j.linalg.activations.impl.ActivationReLU;
import org.nd4j.linalg.activations.impl.ActivationTanH;

import java.util.ArrayList;
import java.util.List;

/**
 * Walks the inner {@link ComputationGraph} of an OdeVertex in topological order (the order ForwardPass.java:62-102
 * evaluates it in) and describes it as the {@code n4j_layer_desc[]} chain of include/node4j.h.
 * <p>
 * The native set is: BatchNormalization, ConvolutionLayer (3x3, Same, stride 1, no bias), DenseLayer, ActivationLayer,
 * activations identity / ReLU / ELU / tanh, and — for time dependent f — a first vertex ShapeMatchVertex(MergeVertex)
 * (ode/vertex/conf/OdeVertex.java:251). Anything else throws {@link UnsupportedOperationException}: there is no CPU
 * fallback inside the native path (the conf classes decide up front whether to take it, see {@link NativeOdeHelper#supports}).
 * <p>
 * NOT COMPILED IN THIS REPOSITORY (no JDK / DL4J in the build image).
 */
final class NativeGraphDescription {

    static final class LayerDesc {
        final int kind;
        final int activation;
        final long nIn;
        final long nOut;
        final boolean hasBias;
        final float eps;

        LayerDesc(int kind, int activation, long nIn, long nOut, boolean hasBias, float eps) {
            this.kind = kind;
            this.activation = activation;
            this.nIn = nIn;
            this.nOut = nOut;
            this.hasBias = hasBias;
            this.eps = eps;
        }
    }

    final List<LayerDesc> layers = new ArrayList<>();

    /** @return null when the graph is outside the native set (reason in {@code why[0]}), else its description */
    static NativeGraphDescription tryDescribe(ComputationGraph graph, long[] stateShape, String[] why) {
        try {
            return new NativeGraphDescription(graph, stateShape);
        } catch (UnsupportedOperationException e) {
            if (why != null && why.length > 0) why[0] = e.getMessage();
            return null;
        }
    }

    NativeGraphDescription(ComputationGraph graph, long[] stateShape) {
        final GraphVertex[] vertices = graph.getVertices();
        String previous = null;
        for (int idx : graph.topologicalSortOrder()) {
            final GraphVertex v = vertices[idx];
            if (v.isInputVertex()) {
                continue;
            }
            // a chain: every vertex consumes exactly the vertex in front of it (plus the time for the first one)
            if (v.getNumInputArrays() > (v instanceof ShapeMatchVertex ? 2 : 1)) {
                throw new UnsupportedOperationException("native ODE helper: vertex " + v.getVertexName() + " has " + v.getNumInputArrays()
                        + " inputs; only chains of layers are supported");
            }
            if (v instanceof ShapeMatchVertex) {
                if (previous != null) {
                    throw new UnsupportedOperationException("native ODE helper: time merge must be the first vertex, found it after " + previous);
                }
                // n_out = number of channels / features the duplicated time occupies (DuplicateScalarToShape.java:41-60): 1 unless overridden
                layers.add(new LayerDesc(Node4j.LAYER_TIME_CONCAT, Node4j.ACT_IDENTITY, 0, ((ShapeMatchVertex) v).timeChannels(stateShape[1]), false, 0f));
            } else if (v instanceof LayerVertex) {
                layers.add(describe(((LayerVertex) v).getLayer(), v.getVertexName()));
            } else {
                throw new UnsupportedOperationException("native ODE helper: vertex type " + v.getClass().getName() + " (" + v.getVertexName() + ")");
            }
            previous = v.getVertexName();
        }
        if (layers.isEmpty()) {
            throw new UnsupportedOperationException("native ODE helper: empty inner graph");
        }
    }

    private static LayerDesc describe(Layer layer, String name) {
        final org.deeplearning4j.nn.conf.layers.Layer conf = layer.conf().getLayer();
        if (conf instanceof BatchNormalization) {
            final BatchNormalization bn = (BatchNormalization) conf;
            if (bn.isLockGammaBeta()) {
                throw new UnsupportedOperationException("native ODE helper: lockGammaBeta BatchNormalization (" + name + ")");
            }
            return new LayerDesc(Node4j.LAYER_BATCHNORM, activation(bn.getActivationFn(), name), bn.getNIn(), bn.getNOut(), false, (float) bn.getEps());
        }
        if (conf instanceof ConvolutionLayer) {
            final ConvolutionLayer c = (ConvolutionLayer) conf;
            final boolean k3 = c.getKernelSize()[0] == 3 && c.getKernelSize()[1] == 3;
            final boolean s1 = c.getStride()[0] == 1 && c.getStride()[1] == 1;
            final boolean d1 = c.getDilation()[0] == 1 && c.getDilation()[1] == 1;
            if (!k3 || !s1 || !d1 || c.getConvolutionMode() != ConvolutionMode.Same || c.hasBias()) {
                throw new UnsupportedOperationException("native ODE helper: convolution " + name + " must be 3x3, stride 1, Same, no bias "
                        + "(examples/mnist/LayerUtil.java:65-72)");
            }
            return new LayerDesc(Node4j.LAYER_CONV3X3, activation(c.getActivationFn(), name), c.getNIn(), c.getNOut(), false, 0f);
        }
        if (conf instanceof DenseLayer) {
            final DenseLayer d = (DenseLayer) conf;
            return new LayerDesc(Node4j.LAYER_DENSE, activation(d.getActivationFn(), name), d.getNIn(), d.getNOut(), d.hasBias(), 0f);
        }
        if (conf instanceof ActivationLayer) {
            return new LayerDesc(Node4j.LAYER_ACTIVATION, activation(((ActivationLayer) conf).getActivationFn(), name), 0, 0, false, 0f);
        }
        throw new UnsupportedOperationException("native ODE helper: layer type " + conf.getClass().getName() + " (" + name + ")");
    }

    private static int activation(IActivation a, String name) {
        if (a == null || a instanceof ActivationIdentity) return Node4j.ACT_IDENTITY;
        if (a instanceof ActivationReLU) return Node4j.ACT_RELU;
        if (a instanceof ActivationELU) return Node4j.ACT_ELU;   // alpha = 1 (the DL4J default; examples/spiral/LayerUtil.java:50-78)
        if (a instanceof ActivationTanH) return Node4j.ACT_TANH;
        throw new UnsupportedOperationException("native ODE helper: activation " + a + " (" + name + ")");
    }
}
