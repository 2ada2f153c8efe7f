//! `extern "C"` declarations of include/ws_b200.h -- only what the shim calls.
//! Every struct is `#[repr(C)]` with the field order of the header.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

pub const WS_OK: c_int = 0;
pub const WS_ERR_MAX_TOO_HIGH: c_int = 2;
pub const WS_ERR_MAX_TOO_LOW: c_int = 3;

pub const WS_SEGMENTING: u8 = 0;
pub const WS_MERGING: u8 = 1;
pub const WS_TIE_FIRST: u8 = 0;
pub const WS_TIE_RANDOM: u8 = 1;

// ws_dtype
pub const WS_F32: c_int = 0;
pub const WS_F64: c_int = 1;
pub const WS_I32: c_int = 2;
pub const WS_U16: c_int = 3;
pub const WS_I16: c_int = 4;
pub const WS_U8: c_int = 5;
pub const WS_I64: c_int = 6;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ws_config {
  pub kind: u8,
  pub max_water_level: u8,
  pub edge_correction: u8,
  pub tie_break: u8,
}

#[repr(C)]
pub struct ws_image {
  pub data: *const u8,
  pub rows: usize,
  pub cols: usize,
  pub row_stride: isize,
  pub col_stride: isize,
}

#[repr(C)]
pub struct ws_hook_ctx {
  pub water_level: u8,
  pub max_water_level: u8,
  pub image: *const u8,
  pub colours: *const u64,
  pub rows: usize,
  pub cols: usize,
  pub seeds: *const u64, // [nseeds][3] = (colour, row, col)
  pub nseeds: usize,
}

#[repr(C)]
pub struct ws_ctx {
  _private: [u8; 0],
}

pub type ws_level_hook = extern "C" fn(user: *mut c_void, hctx: *const ws_hook_ctx);

extern "C" {
  pub fn ws_ctx_create(device: c_int, out: *mut *mut ws_ctx) -> c_int;
  pub fn ws_ctx_destroy(ctx: *mut ws_ctx);
  pub fn ws_last_error(ctx: *const ws_ctx) -> *const c_char;
  pub fn ws_status_str(status: c_int) -> *const c_char;
  pub fn ws_abi_version() -> c_int;
  pub fn ws_free(p: *mut c_void);
  pub fn ws_ctx_set_tie_seed(ctx: *mut ws_ctx, seed: u64) -> c_int;
  pub fn ws_config_validate(cfg: *const ws_config) -> c_int;
  pub fn ws_output_shape(
    cfg: *const ws_config,
    rows: usize,
    cols: usize,
    out_rows: *mut usize,
    out_cols: *mut usize,
  ) -> c_int;
  pub fn ws_find_local_minima(
    ctx: *mut ws_ctx,
    img: *const ws_image,
    out_rc: *mut *mut u64,
    out_n: *mut usize,
  ) -> c_int;
  pub fn ws_transform(
    ctx: *mut ws_ctx,
    cfg: *const ws_config,
    img: *const ws_image,
    seeds_rc: *const u64,
    nseeds: usize,
    out_labels: *mut u64,
  ) -> c_int;
  pub fn ws_transform_history(
    ctx: *mut ws_ctx,
    cfg: *const ws_config,
    img: *const ws_image,
    seeds_rc: *const u64,
    nseeds: usize,
    out_levels: *mut u8,
    out_labels: *mut u64,
  ) -> c_int;
  pub fn ws_transform_to_list(
    ctx: *mut ws_ctx,
    cfg: *const ws_config,
    img: *const ws_image,
    seeds_rc: *const u64,
    nseeds: usize,
    out_levels: *mut u8,
    out_sizes: *mut u64,
  ) -> c_int;
  pub fn ws_transform_with_hook(
    ctx: *mut ws_ctx,
    cfg: *const ws_config,
    img: *const ws_image,
    seeds_rc: *const u64,
    nseeds: usize,
    hook: Option<ws_level_hook>,
    user: *mut c_void,
  ) -> c_int;
  pub fn ws_pre_processor(
    ctx: *mut ws_ctx,
    dtype: c_int,
    data: *const c_void,
    n: usize,
    max_value: u8,
    out: *mut u8,
  ) -> c_int;
}
