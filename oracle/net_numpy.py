"""TEST INFRASTRUCTURE ONLY.  Independent numpy (float64) restatement of the reference's TF1 graph
(policy_value_network.py:45-74, 151-162) in TensorFlow's own conventions: NHWC activations, conv kernels [kh, kw, cin, cout]
with 'SAME' padding, tf.contrib.layers.batch_norm(center=False, scale=False, epsilon=1e-5) in inference mode
((x - moving_mean) / sqrt(moving_var + eps)), tf.reshape flattening in (h, w, c) order, fully_connected weights [in, out].
NN parity with the reference itself stays UNPINNED (no TensorFlow, no checkpoint: SURVEY 0.8); this file only cross-checks
the PyTorch re-implementation (cchess_zero_b200/net.py) against a second, differently written evaluation of the same graph."""
import numpy as np


def conv2d_same(x, k, b):
    """x [B,H,W,Cin], k [kh,kw,Cin,Cout], stride 1, SAME padding (tf.layers.conv2d)."""
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = k.shape
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xp = np.zeros((B, H + kh - 1, W + kw - 1, Cin), dtype=np.float64)
    xp[:, ph:ph + H, pw:pw + W, :] = x
    out = np.zeros((B, H, W, Cout), dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            out += xp[:, i:i + H, j:j + W, :] @ k[i, j]
    return out + b


def bn_inference(x, mean, var, eps=1e-5):
    return (x - mean) / np.sqrt(var + eps)


def relu(x):
    return np.maximum(x, 0.0)


def forward(x, p, n_blocks):
    """x [B,9,10,14]; p: dict of TF-layout float64 arrays (see tf_params_from_torch).  -> (logits [B,2086], value [B,1])"""
    h = relu(bn_inference(conv2d_same(x, p["conv_in.k"], p["conv_in.b"]), p["bn_in.mean"], p["bn_in.var"]))
    for i in range(n_blocks):                                     # residual_block, policy_value_network.py:151-162
        y = relu(bn_inference(conv2d_same(h, p["b%d.c1.k" % i], p["b%d.c1.b" % i]), p["b%d.b1.mean" % i], p["b%d.b1.var" % i]))
        y = bn_inference(conv2d_same(y, p["b%d.c2.k" % i], p["b%d.c2.b" % i]), p["b%d.b2.mean" % i], p["b%d.b2.var" % i])
        h = relu(h + y)
    ph = relu(bn_inference(conv2d_same(h, p["p_conv.k"], p["p_conv.b"]), p["p_bn.mean"], p["p_bn.var"]))
    logits = ph.reshape(len(x), 9 * 10 * 2) @ p["p_fc.w"] + p["p_fc.b"]          # tf.reshape [-1, 180]; no softmax (line 64)
    vh = relu(bn_inference(conv2d_same(h, p["v_conv.k"], p["v_conv.b"]), p["v_bn.mean"], p["v_bn.var"]))
    v = relu(vh.reshape(len(x), 90) @ p["v_fc1.w"] + p["v_fc1.b"])
    value = np.tanh(v @ p["v_fc2.w"] + p["v_fc2.b"])
    return logits, value


def tf_params_from_torch(net):
    """Re-layout a cchess_zero_b200.net.PolicyValueNet's parameters the way TensorFlow stores them."""
    def k(conv):
        return conv.weight.detach().double().permute(2, 3, 1, 0).numpy(), conv.bias.detach().double().numpy()

    def bn(b):
        return b.running_mean.double().numpy(), b.running_var.double().numpy()

    def fc(l):
        return l.weight.detach().double().t().numpy(), l.bias.detach().double().numpy()
    p = {}
    p["conv_in.k"], p["conv_in.b"] = k(net.conv_in)
    p["bn_in.mean"], p["bn_in.var"] = bn(net.bn_in)
    for i, blk in enumerate(net.blocks):
        p["b%d.c1.k" % i], p["b%d.c1.b" % i] = k(blk.c1)
        p["b%d.b1.mean" % i], p["b%d.b1.var" % i] = bn(blk.b1)
        p["b%d.c2.k" % i], p["b%d.c2.b" % i] = k(blk.c2)
        p["b%d.b2.mean" % i], p["b%d.b2.var" % i] = bn(blk.b2)
    p["p_conv.k"], p["p_conv.b"] = k(net.p_conv)
    p["p_bn.mean"], p["p_bn.var"] = bn(net.p_bn)
    p["p_fc.w"], p["p_fc.b"] = fc(net.p_fc)
    p["v_conv.k"], p["v_conv.b"] = k(net.v_conv)
    p["v_bn.mean"], p["v_bn.var"] = bn(net.v_bn)
    p["v_fc1.w"], p["v_fc1.b"] = fc(net.v_fc1)
    p["v_fc2.w"], p["v_fc2.b"] = fc(net.v_fc2)
    return p
