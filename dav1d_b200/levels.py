"""Enumerations shared with dav1d (numbering is the reference's: src/levels.h:38-110)."""

(TX_4X4, TX_8X8, TX_16X16, TX_32X32, TX_64X64, RTX_4X8, RTX_8X4, RTX_8X16, RTX_16X8, RTX_16X32,
 RTX_32X16, RTX_32X64, RTX_64X32, RTX_4X16, RTX_16X4, RTX_8X32, RTX_32X8, RTX_16X64,
 RTX_64X16) = range(19)
N_RECT_TX_SIZES = 19

(DCT_DCT, ADST_DCT, DCT_ADST, ADST_ADST, FLIPADST_DCT, DCT_FLIPADST, FLIPADST_FLIPADST,
 ADST_FLIPADST, FLIPADST_ADST, IDTX, V_DCT, H_DCT, V_ADST, H_ADST, V_FLIPADST, H_FLIPADST,
 WHT_WHT) = range(17)
N_TX_TYPES = 16
N_TX_TYPES_PLUS_LL = 17

TX_NAMES = ["4x4", "8x8", "16x16", "32x32", "64x64", "4x8", "8x4", "8x16", "16x8", "16x32", "32x16",
            "32x64", "64x32", "4x16", "16x4", "8x32", "32x8", "16x64", "64x16"]
TXTP_NAMES = ["dct_dct", "adst_dct", "dct_adst", "adst_adst", "flipadst_dct", "dct_flipadst",
              "flipadst_flipadst", "adst_flipadst", "flipadst_adst", "idtx", "v_dct", "h_dct", "v_adst",
              "h_adst", "v_flipadst", "h_flipadst", "wht_wht"]
# transform width / height in pixels per RectTxfmSize (reference src/tables.c:129-168)
TX_W = [4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64]
TX_H = [4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16]


def tx_coef_dims(tx):
    """(sw, sh): dimensions of the coded coefficient block, min(w,32) x min(h,32)."""
    return min(TX_W[tx], 32), min(TX_H[tx], 32)


def itx_defined(tx, txtp):
    """Which itxfm_add[tx][txtp] slots dav1d fills (reference src/itx_tmpl.c:220-288)."""
    if txtp == WHT_WHT:
        return tx == TX_4X4
    mx, mn = max(TX_W[tx], TX_H[tx]), min(TX_W[tx], TX_H[tx])
    if mx == 64:
        return txtp == DCT_DCT
    if mx == 32:
        return txtp in (DCT_DCT, IDTX)
    if mx == 16 and mn == 16:
        return txtp <= H_DCT
    return True
