"""CPU restatement (test infrastructure) of the LOCAL-ROWS formulation of the symmetric InfoNCE
loss that the CUDA kernels `b200clip_clip_loss_fwd/bwd` implement for data parallelism.

Reference semantics: CLIP/train.py:162-166 on one device
    loss = (CE(logits_per_image, arange) + CE(logits_per_text, arange)) / 2 .
Sharded over R ranks, rank r owns rows [r*Bl, (r+1)*Bl) of the global batch and
  forward : lse_i[i] = logsumexp_j s<img_i, txt_j>,  lse_t[i] = logsumexp_j s<txt_i, img_j>  (local i, all j)
            loss = sum over ranks of sum_i (lse_i[i] + lse_t[i] - 2 s<img_i,txt_i>) / (2 Bg)
  backward: with the all-gathered lse vectors, for local i
            d_img[i] = s/(2Bg) sum_j (p_img[i,j] + p_txt[j,i] - 2 delta_ij) txt_j
            d_txt[i] = s/(2Bg) sum_j (p_txt[i,j] + p_img[j,i] - 2 delta_ij) img_j
            d_logit_scale (local share) = s/(2Bg) sum_{local i, all j} G[i,j] <img_i, txt_j>
which is the exact gradient of the GLOBAL loss w.r.t. the local embeddings: no gradient collective.
"""
import torch


def local_fwd(img_all, txt_all, logit_scale, row0, bl):
    s = logit_scale.exp()
    li = s * img_all[row0:row0 + bl] @ txt_all.t()        # [bl, Bg]
    lt = s * txt_all[row0:row0 + bl] @ img_all.t()
    lse_i, lse_t = torch.logsumexp(li, 1), torch.logsumexp(lt, 1)
    diag = s * (img_all[row0:row0 + bl] * txt_all[row0:row0 + bl]).sum(1)
    loss_sum = torch.stack([(lse_i - diag).sum(), (lse_t - diag).sum()])
    correct = (li.argmax(1) == torch.arange(row0, row0 + bl)).sum()
    return lse_i, lse_t, loss_sum, correct


def local_bwd(img_all, txt_all, logit_scale, lse_i_all, lse_t_all, row0, bl, grad_out=1.0):
    Bg = img_all.shape[0]
    s = logit_scale.exp()
    coef = grad_out * s / (2.0 * Bg)
    d = img_all[row0:row0 + bl] @ txt_all.t()             # <img_i, txt_j>
    eye = torch.zeros(bl, Bg, dtype=d.dtype)
    eye[torch.arange(bl), torch.arange(row0, row0 + bl)] = 1.0
    g0 = torch.exp(s * d - lse_i_all[row0:row0 + bl, None]) + torch.exp(s * d - lse_t_all[None, :]) - 2 * eye
    d2 = txt_all[row0:row0 + bl] @ img_all.t()            # <txt_i, img_j>
    g1 = torch.exp(s * d2 - lse_t_all[row0:row0 + bl, None]) + torch.exp(s * d2 - lse_i_all[None, :]) - 2 * eye
    return coef * g0 @ txt_all, coef * g1 @ img_all, coef * (g0 * d).sum()
