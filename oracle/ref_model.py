"""oracle/ref_model.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Op-by-op restatement of the reference's Point-Transformer + CBL network and loss
(LiyaoTang/contrastBoundary, pytorch/model/{pointtransformer_seg,blocks,heads,basic_operators}.py)
in plain PyTorch, for (a) parity tests of the fused CUDA path and (b) the CPU baseline legs of
bench.py (`--impl reference`, `cpu_baseline`): /root/reference does not exist on the GPU box, so
the reference's python model cannot be imported there.

It keeps the reference's execution structure on purpose — every neighbour search is recomputed
where the reference recomputes it (blocks.py:34-35), BatchNorm goes through the same
transpose/contiguous round trips (blocks.py:38,40), interpolation is the python loop over k
(pointops.py:176-177), TransitionDown sizes are computed per call — so that its run time is the
reference's run time.  The operator backend `ops` is a module with the reference's pointops API:
oracle.cpu_pointops (CPU, C restatement) or a CUDA one (the reference's own pointops_cuda.so via
oracle.gpu_pointops, "stock pointops" baseline B1).

Pinned by tests/golden/model_ref.npz, which tests/golden/make_golden_model.py produces by importing
the REAL reference model code from /root/reference on CPU (tests/test_model_oracle_cpu.py).
Parameter names equal the reference's, so one state_dict drives reference, oracle and product.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

_EPS = 1e-12  # basic_operators.py:7


def _int_tensor(vals, like):
    return torch.tensor(vals, dtype=torch.int32, device=like.device)


class RefLayer(nn.Module):
    """blocks.py:14-44"""

    def __init__(self, ops, in_planes, out_planes, share_planes=8, nsample=16):
        super().__init__()
        self.ops = ops
        self.mid_planes = mid = out_planes
        self.out_planes, self.share_planes, self.nsample = out_planes, share_planes, nsample
        self.linear_q = nn.Linear(in_planes, mid)
        self.linear_k = nn.Linear(in_planes, mid)
        self.linear_v = nn.Linear(in_planes, out_planes)
        self.linear_p = nn.Sequential(nn.Linear(3, 3), nn.BatchNorm1d(3), nn.ReLU(inplace=True), nn.Linear(3, out_planes))
        self.linear_w = nn.Sequential(nn.BatchNorm1d(mid), nn.ReLU(inplace=True), nn.Linear(mid, mid // share_planes),
                                      nn.BatchNorm1d(mid // share_planes), nn.ReLU(inplace=True),
                                      nn.Linear(out_planes // share_planes, out_planes // share_planes))

    def forward(self, p, x, o):
        q, k, v = self.linear_q(x), self.linear_k(x), self.linear_v(x)
        kg = self.ops.queryandgroup(self.nsample, p, p, k, None, o, o, use_xyz=True)      # :34
        vg = self.ops.queryandgroup(self.nsample, p, p, v, None, o, o, use_xyz=False)     # :35 (search repeated)
        pr, kg = kg[:, :, 0:3], kg[:, :, 3:]
        for i, layer in enumerate(self.linear_p):                                          # :38
            pr = layer(pr.transpose(1, 2).contiguous()).transpose(1, 2).contiguous() if i == 1 else layer(pr)
        w = kg - q.unsqueeze(1) + pr                                                        # :39 (out == mid)
        for i, layer in enumerate(self.linear_w):                                          # :40
            w = layer(w.transpose(1, 2).contiguous()).transpose(1, 2).contiguous() if i % 3 == 0 else layer(w)
        w = F.softmax(w, dim=1)                                                             # :41
        n, ns, c = vg.shape
        s = self.share_planes
        return ((vg + pr).view(n, ns, s, c // s) * w.unsqueeze(2)).sum(1).view(n, c)       # :43


class RefDown(nn.Module):
    """blocks.py:47-77"""

    def __init__(self, ops, in_planes, out_planes, stride=1, nsample=16):
        super().__init__()
        self.ops, self.stride, self.nsample = ops, stride, nsample
        self.linear = nn.Linear((3 if stride != 1 else 0) + in_planes, out_planes, bias=False)
        self.bn = nn.BatchNorm1d(out_planes)

    def forward(self, pxo):
        p, x, o = pxo
        if self.stride == 1:
            return [p, F.relu(self.bn(self.linear(x))), o]
        oh = o.tolist()                                                                     # host sync, as :64-66
        n_o, count, prev = [], 0, 0
        for e in oh:
            count += (e - prev) // self.stride
            n_o.append(count)
            prev = e
        n_o = _int_tensor(n_o, o)
        idx = self.ops.furthestsampling(p, o, n_o)                                          # :69
        n_p = p[idx.long(), :]
        x = self.ops.queryandgroup(self.nsample, p, n_p, x, None, o, n_o, use_xyz=True)     # :71
        x = F.relu(self.bn(self.linear(x).transpose(1, 2).contiguous()))                    # :72
        x = F.max_pool1d(x, self.nsample).squeeze(-1)                                       # :73
        return [n_p, x, n_o]


class RefUp(nn.Module):
    """blocks.py:80-109"""

    def __init__(self, ops, in_planes, out_planes=None):
        super().__init__()
        self.ops = ops
        if out_planes is None:
            self.linear1 = nn.Sequential(nn.Linear(2 * in_planes, in_planes), nn.BatchNorm1d(in_planes), nn.ReLU(inplace=True))
            self.linear2 = nn.Sequential(nn.Linear(in_planes, in_planes), nn.ReLU(inplace=True))
        else:
            self.linear1 = nn.Sequential(nn.Linear(out_planes, out_planes), nn.BatchNorm1d(out_planes), nn.ReLU(inplace=True))
            self.linear2 = nn.Sequential(nn.Linear(in_planes, out_planes), nn.BatchNorm1d(out_planes), nn.ReLU(inplace=True))

    def forward(self, pxo1, pxo2=None):
        if pxo2 is None:
            _, x, o = pxo1
            oh = o.tolist()
            parts, prev = [], 0
            for e in oh:                                                                    # :94-103
                xb = x[prev:e, :]
                cnt = e - prev
                parts.append(torch.cat((xb, self.linear2(xb.sum(0, True) / cnt).repeat(cnt, 1)), 1))
                prev = e
            return self.linear1(torch.cat(parts, 0))
        p1, x1, o1 = pxo1
        p2, x2, o2 = pxo2
        return self.linear1(x1) + self.ops.interpolation(p2, p1, self.linear2(x2), o2, o1)   # :108


class RefBlock(nn.Module):
    """blocks.py:112-133"""

    def __init__(self, ops, in_planes, planes, share_planes=8, nsample=16):
        super().__init__()
        self.linear1 = nn.Linear(in_planes, planes, bias=False)
        self.bn1 = nn.BatchNorm1d(planes)
        self.transformer2 = RefLayer(ops, planes, planes, share_planes, nsample)
        self.bn2 = nn.BatchNorm1d(planes)
        self.linear3 = nn.Linear(planes, planes, bias=False)
        self.bn3 = nn.BatchNorm1d(planes)

    def forward(self, pxo):
        p, x, o = pxo
        identity = x
        x = F.relu(self.bn1(self.linear1(x)))
        x = F.relu(self.bn2(self.transformer2(p, x, o)))
        x = self.bn3(self.linear3(x))
        return [p, F.relu(x + identity), o]


class _RefMLP(nn.Module):
    def __init__(self, fdim, d_out):
        super().__init__()
        self.infer = nn.Sequential(nn.Linear(fdim, d_out), nn.BatchNorm1d(d_out), nn.ReLU(inplace=True))


class RefMultiHead(nn.Module):
    """heads.py:13-61 (stage Ua, latent, concat)"""

    def __init__(self, ops, fdims, base_fdim, classes):
        super().__init__()
        self.ops = ops
        self.infer_list = nn.ModuleList([_RefMLP(f, base_fdim) for f in fdims])
        self.cls = nn.Linear(base_fdim * len(fdims), classes)

    def forward(self, up):
        p0, _, o0 = up[0]["pxo"]
        collect = []
        for i, (st, mlp) in enumerate(zip(up, self.infer_list)):
            p, x, o = st["pxo"]
            lat = mlp.infer(x)
            st["latent"] = lat
            collect.append(lat if i == 0 else self.ops.interpolation(p, p0, lat, o, o0, k=1))   # :44-51
        return self.cls(torch.cat(collect, 1))


class RefSeg(nn.Module):
    """pointtransformer_seg.py:27-143"""

    def __init__(self, ops, c=6, k=13, planes=(32, 64, 128, 256, 512), blocks=(2, 3, 4, 6, 3), share_planes=8,
                 base_fdim=32):
        super().__init__()
        self.ops, self.c, self.in_planes = ops, c, c
        stride, nsample = [1, 4, 4, 4, 4], [8, 16, 16, 16, 16]
        for i in range(5):
            setattr(self, f"enc{i + 1}", self._enc(planes[i], blocks[i], share_planes, stride[i], nsample[i]))
        for i in range(4, -1, -1):
            setattr(self, f"dec{i + 1}", self._dec(planes[i], 2, share_planes, nsample[i], i == 4))
        self.head = RefMultiHead(ops, planes, base_fdim, k)

    def _enc(self, planes, blocks, sp, stride, nsample):
        layers = [RefDown(self.ops, self.in_planes, planes, stride, nsample)]
        self.in_planes = planes
        layers += [RefBlock(self.ops, planes, planes, sp, nsample) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def _dec(self, planes, blocks, sp, nsample, is_head):
        layers = [RefUp(self.ops, self.in_planes, None if is_head else planes)]
        self.in_planes = planes
        layers += [RefBlock(self.ops, planes, planes, sp, nsample) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, inputs):
        p0, x0, o0 = inputs["points"], inputs["features"], inputs["offset"]
        x0 = p0 if self.c == 3 else torch.cat((p0, x0), 1)
        pxo = [p0, x0, o0]
        enc_out = []
        for enc in (self.enc1, self.enc2, self.enc3, self.enc4, self.enc5):
            pxo = enc(pxo)
            enc_out.append(list(pxo))
        (p1, x1, o1), (p2, x2, o2), (p3, x3, o3), (p4, x4, o4), (p5, x5, o5) = enc_out
        x5 = self.dec5[1:]([p5, self.dec5[0]([p5, x5, o5]), o5])[1]
        x4 = self.dec4[1:]([p4, self.dec4[0]([p4, x4, o4], [p5, x5, o5]), o4])[1]
        x3 = self.dec3[1:]([p3, self.dec3[0]([p3, x3, o3], [p4, x4, o4]), o3])[1]
        x2 = self.dec2[1:]([p2, self.dec2[0]([p2, x2, o2], [p3, x3, o3]), o2])[1]
        x1 = self.dec1[1:]([p1, self.dec1[0]([p1, x1, o1], [p2, x2, o2]), o1])[1]
        up = [{"pxo": (p1, x1, o1)}, {"pxo": (p2, x2, o2)}, {"pxo": (p3, x3, o3)}, {"pxo": (p4, x4, o4)},
              {"pxo": (p5, x5, o5)}]
        return self.head(up), up


class RefLoss(nn.Module):
    """pointtransformer_seg.py:15-25 + heads.py:185-253 + basic_operators.py:9-50 (softnn/l2/cnt/T=1/w.1)"""

    def __init__(self, ops, classes=13, nsample=(36, 24, 24, 24, 24), nstride=(4, 4, 4, 4), temperature=1.0,
                 weight=0.1, ignore_label=255):
        super().__init__()
        self.ops, self.classes, self.nsample, self.nstride = ops, classes, list(nsample), list(nstride)
        self.temperature, self.weight = temperature, weight
        self.xen = nn.CrossEntropyLoss(ignore_index=ignore_label)

    def stage(self, i, up, target):
        p, _, o = up[i]["pxo"]
        feat = up[i]["latent"]
        labels = F.one_hot(target, self.classes)
        if i == 0:
            labels = labels.float()
        else:
            kr = 1
            for s in self.nstride[:i]:
                kr *= s
            p0, _, o0 = up[0]["pxo"]
            nidx, _ = self.ops.knnquery(kr, p0, p, o0, o)
            labels = labels[nidx.view(-1).long(), :].view(p.shape[0], kr, self.classes).float().mean(-2)
        ns = self.nsample[i]
        idx, _ = self.ops.knnquery(ns, p, p, o, o)
        ns -= 1
        idx = idx[..., 1:].contiguous()
        m = idx.shape[0]
        nb_label = labels[idx.view(-1).long(), :].view(m, ns, self.classes)
        nb_feat = feat[idx.view(-1).long(), :].view(m, ns, feat.shape[1])
        posmask = torch.argmax(labels.unsqueeze(-2), -1) == torch.argmax(nb_label, -1)
        pm = posmask.int().sum(-1)
        pm = torch.logical_and(0 < pm, pm < ns)
        if not torch.any(pm):
            return torch.zeros((), device=feat.device)
        posmask, f, nf = posmask[pm], feat[pm], nb_feat[pm]
        dist = torch.sqrt(torch.sum((f.unsqueeze(-2) - nf) ** 2, -1) + _EPS)
        d = -dist
        d = d - torch.max(d, -1, keepdim=True)[0]
        if self.temperature is not None:
            d = d / self.temperature
        e = torch.exp(d)
        loss = -torch.log(torch.sum(e * posmask, -1) / torch.sum(e, -1) + _EPS)
        return torch.mean(loss) * float(self.weight)

    def forward(self, output, target, up):
        return torch.stack([self.xen(output, target)] + [self.stage(i, up, target) for i in range(len(up))])
