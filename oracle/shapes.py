"""TEST INFRASTRUCTURE ONLY -- state-dict key -> shape tables of the reference Generator / Discriminator, derived from the
config alone (constructors: networks_epigraf.py:76-112,134-189,266-283; networks_stylegan2.py:94-126,156-166,181-229;
networks_discriminator.py:19-65,129-155,188-254; layers.py:22-40,66-125,182-226,251-275).  Lets the CPU arm of bench.py
build random-init weights for oracle/restated.py without touching the product package."""
import math


def _fc(d, p, i, o, bias=True):
    d[p + 'weight'] = (o, i)
    if bias:
        d[p + 'bias'] = (o,)


def _mapping(d, p, z_dim, c_dim, w_dim, num_ws, layers=2):
    emb = 0
    if c_dim > 0:
        _fc(d, p + 'embed.', c_dim, w_dim); emb = w_dim
    feats = [z_dim + emb] + [w_dim] * layers
    for i in range(layers):
        _fc(d, f'{p}fc{i}.', feats[i], feats[i + 1])
    if num_ws is not None:
        d[p + 'w_avg'] = (w_dim,)


def generator_shapes(g):
    d = {}
    tri = g['tri_plane']
    res_list = [2 ** i for i in range(2, int(math.log2(tri['res'])) + 1)]
    ch = {r: min(int(g['cbase'] * g['fmaps']) // r, g['cmax']) for r in res_list}
    oc = 3 * tri['feat_dim']; w = g['w_dim']
    num_ws = 0
    for r in res_list:
        p = f'synthesis.tri_plane_decoder.b{r}.'
        d[p + 'resample_filter'] = (4, 4)
        convs = ['conv1'] if r == 4 else ['conv0', 'conv1']
        if r == 4:
            d[p + 'const'] = (ch[r], r, r)
        for cv in convs:
            cin = ch[r // 2] if cv == 'conv0' else ch[r]
            q = p + cv + '.'
            d[q + 'weight'] = (ch[r], cin, 3, 3); d[q + 'noise_strength'] = (); d[q + 'bias'] = (ch[r],)
            d[q + 'resample_filter'] = (4, 4); d[q + 'noise_const'] = (r, r)
            _fc(d, q + 'affine.', w, cin)
            num_ws += 1
        q = p + 'torgb.'
        d[q + 'weight'] = (oc, ch[r], 1, 1); d[q + 'bias'] = (oc,); _fc(d, q + 'affine.', w, ch[r])
    num_ws += 1
    _fc(d, 'synthesis.tri_plane_mlp.model.0.', tri['feat_dim'], tri['mlp']['hid_dim'])
    _fc(d, 'synthesis.tri_plane_mlp.model.1.', tri['mlp']['hid_dim'], 4)
    da = g['depth_adaptor']
    if da['enabled']:
        dims = [1] + [da['hid_dim']] * da['num_hid_layers']
        for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
            q = f'synthesis.depth_adaptor.layers.{i}.'
            d[q + 'weight'] = (b, a, da['kernel_size'], da['kernel_size']); d[q + 'bias'] = (b,); d[q + 'resample_filter'] = (4, 4)
        d['synthesis.depth_adaptor.head.weight'] = (1, dims[-1], 1, 1); d['synthesis.depth_adaptor.head.bias'] = (1,)
        d['synthesis.depth_adaptor.head.resample_filter'] = (4, 4)
        d['synthesis.depth_adaptor.progress_coef'] = (1,); d['synthesis.depth_adaptor.near_plane_offset_raw'] = (1,)
    _mapping(d, 'mapping.', g['z_dim'], g['c_dim'], w, num_ws, g['map_depth'])
    return d, num_ws


def discriminator_shapes(dc, patch_res, img_channels=4, feat_predict_dim=0):
    d = {}
    top = patch_res * 2 ** dc['num_additional_start_blocks']
    res_list = [2 ** i for i in range(int(math.log2(top)), 2, -1)]
    ch = {r: min(int(dc['cbase'] * dc['fmaps']) // r, dc['cmax']) for r in res_list + [4]}
    enc_dim = 3 * (2 * 10 + 256)                         # ScalarEncoder1d(coord_dim=3, x_multiplier=1000, const_emb_dim=256)
    d['scalar_enc.const_embed.weight'] = (1001, 256); d['scalar_enc.fourier_encoder.fourier_coefs'] = (10,)
    _mapping(d, 'hyper_mod_mapping.', 0, enc_dim, 512, None)
    for i, r in enumerate(res_list):
        p = f'b{r}.'
        d[p + 'resample_filter'] = (4, 4)
        for nm, (o, c_, k, b) in dict(fromrgb=(ch[r], img_channels, 1, True), conv0=(ch[r], ch[r], 3, True),
                                      conv1=(ch[r // 2], ch[r], 3, True), skip=(ch[r // 2], ch[r], 1, False)).items():
            q = p + nm + '.'
            d[q + 'weight'] = (o, c_, k, k); d[q + 'resample_filter'] = (4, 4)
            if b:
                d[q + 'bias'] = (o,)
        _fc(d, p + 'conv1.affine.', 512, ch[r])
    _mapping(d, 'head_mapping.', 0, dc['c_dim'] + enc_dim, ch[4], None)
    d['b4.conv.weight'] = (ch[4], ch[4] + 1, 3, 3); d['b4.conv.bias'] = (ch[4],); d['b4.conv.resample_filter'] = (4, 4)
    _fc(d, 'b4.fc.', ch[4] * 16, ch[4]); _fc(d, 'b4.out.', ch[4], ch[4])
    if feat_predict_dim > 0:
        _fc(d, 'b4.feat_out.0.', ch[4] * 16, ch[4]); _fc(d, 'b4.feat_out.1.', ch[4], feat_predict_dim)
    return d, res_list
