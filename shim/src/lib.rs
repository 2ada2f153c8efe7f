//! rustronomy-watershed's public API (v0.4.1) over the B200-native CUDA engine.
//!
//! Every public item below has the name, the signature and the behaviour of the item of the reference crate
//! cited next to it (`src/lib.rs` of smups/rustronomy-watershed); only the bodies differ: they call
//! `libws_b200.so` through the C ABI of `include/ws_b200.h` (see `ffi.rs`).  There is no CPU fallback: the
//! first call on a thread creates a context on CUDA device `WS_B200_DEVICE` (default 0) and panics if that fails.
//!
//! Behaviour a user of the reference should know (all stated in ws_b200.h too):
//! * ties between differently coloured neighbours resolve to the first coloured neighbour in the order down,
//!   right, left, up (the reference draws one with `thread_rng`, lib.rs:250-253) -- always one of the reference's
//!   possible outcomes; `TransformBuilder::set_random_tie_break` restores the random draw;
//! * merged lakes carry their smallest seed colour (the reference's choice, lib.rs:539, depends on an
//!   unspecified sort order): equal up to a renumbering per level; counts and sizes are exact;
//! * `SegmentingWatershed::transform` returns the colours at the last level (the reference panics, lib.rs:1821).
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image): `tests/c_abi_smoke.c` drives the same
//! entry points with the same structs from C, and the ctypes binding drives them in every GPU test.
#![allow(clippy::type_complexity)]

#[cfg(not(target_pointer_width = "64"))]
compile_error!("the engine's labels are 64-bit: usize must be u64");

pub mod ffi;
#[cfg(feature = "plots")]
pub mod plotting;

use ndarray as nd;
use num_traits::{Num, ToPrimitive};
use std::os::raw::{c_int, c_void};

// lib.rs:138-141
pub const UNCOLOURED: usize = 0;
pub const NORMAL_MAX: u8 = u8::MAX - 1;
pub const ALWAYS_FILL: u8 = u8::MIN;
pub const NEVER_FILL: u8 = u8::MAX;

/// lib.rs:144-154
pub mod prelude {
  pub use crate::{MergingWatershed, TransformBuilder, Watershed, WatershedUtils};
  #[cfg(feature = "plots")]
  pub mod color_maps {
    pub use crate::plotting::{grey_scale, inferno, magma, plasma, viridis};
  }
}

/// `fn(count, min, max)` of the `plots` feature (lib.rs:913-916)
#[cfg(feature = "plots")]
pub type ColourMap = fn(usize, usize, usize) -> Result<plotting::RGBColor, Box<dyn std::error::Error>>;

/// The two options of the `plots` feature (lib.rs:910-916, 1299-1304); empty without it.
#[derive(Clone)]
struct PlotCfg {
  #[cfg(feature = "plots")]
  path: Option<std::path::PathBuf>,
  #[cfg(feature = "plots")]
  map: Option<ColourMap>,
}

impl PlotCfg {
  const fn none() -> Self {
    PlotCfg {
      #[cfg(feature = "plots")]
      path: None,
      #[cfg(feature = "plots")]
      map: None,
    }
  }
}

////////////////////////////////////////////////////////////////////////////////
//                       context, views, error handling                       //
////////////////////////////////////////////////////////////////////////////////

/// One engine context per thread: the structs stay `Send + Sync` and every method keeps `&self`
/// (a `ws_ctx` is used by one thread at a time, ws_b200.h).
struct Ctx(*mut ffi::ws_ctx);

impl Drop for Ctx {
  fn drop(&mut self) {
    unsafe { ffi::ws_ctx_destroy(self.0) }
  }
}

thread_local! {
  static CTX: Ctx = {
    let device: c_int = std::env::var("WS_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
    let mut raw: *mut ffi::ws_ctx = std::ptr::null_mut();
    let status = unsafe { ffi::ws_ctx_create(device, &mut raw) };
    if status != ffi::WS_OK {
      let msg = unsafe { std::ffi::CStr::from_ptr(ffi::ws_status_str(status)) }.to_string_lossy().into_owned();
      panic!("watershed engine: cannot create a context on CUDA device {device}: {msg} (there is no CPU fallback)");
    }
    Ctx(raw)
  };
}

fn with_ctx<R>(f: impl FnOnce(*mut ffi::ws_ctx) -> R) -> R {
  CTX.with(|c| f(c.0))
}

/// The reference's run-time failures are panics (e.g. an out-of-bounds seed, lib.rs:1366): so are ours.
fn check(ctx: *mut ffi::ws_ctx, status: c_int) {
  if status != ffi::WS_OK {
    let msg = unsafe { std::ffi::CStr::from_ptr(ffi::ws_last_error(ctx)) }.to_string_lossy().into_owned();
    panic!("watershed engine: status {status}: {msg}");
  }
}

/// `ArrayView2<u8>` may be strided (a slice of a cube, a transposed view): the strides go through.
fn view(img: &nd::ArrayView2<u8>) -> ffi::ws_image {
  let s = img.strides();
  ffi::ws_image { data: img.as_ptr(), rows: img.nrows(), cols: img.ncols(), row_stride: s[0], col_stride: s[1] }
}

/// The layout of a Rust tuple is unspecified: repack `&[(usize, usize)]` as `[n][2]` u64.
fn seeds_rc(seeds: &[(usize, usize)]) -> Vec<u64> {
  let mut out = Vec::with_capacity(2 * seeds.len());
  for &(r, c) in seeds {
    out.push(r as u64);
    out.push(c as u64);
  }
  out
}

fn out_shape(edge_correction: bool, input: &nd::ArrayView2<u8>) -> (usize, usize) {
  let pad = if edge_correction { 2 } else { 0 }; // lib.rs:1330-1337: the padding is NOT removed from the results
  (input.nrows() + pad, input.ncols() + pad)
}

////////////////////////////////////////////////////////////////////////////////
//                          WATERSHED TRANSFORMS                              //
////////////////////////////////////////////////////////////////////////////////

/// lib.rs:844-850
#[derive(Clone)]
pub struct HookCtx<'a> {
  pub water_level: u8,
  pub max_water_level: u8,
  pub image: nd::ArrayView2<'a, u8>,
  pub colours: nd::ArrayView2<'a, usize>,
  pub seeds: &'a [(usize, (usize, usize))],
}

/// lib.rs:908-1047, with the two options of the `plots` feature (lib.rs:972-994) behind the same feature gate
#[derive(Clone)]
pub struct TransformBuilder<T = ()> {
  max_water_level: u8,
  edge_correction: bool,
  wlvl_hook: Option<fn(HookCtx) -> T>,
  tie_break: u8,
  tie_seed: Option<u64>,
  plots: PlotCfg,
}

impl Default for TransformBuilder<()> {
  fn default() -> Self {
    TransformBuilder::new()
  }
}

impl<T> TransformBuilder<T> {
  pub const fn new() -> Self {
    TransformBuilder {
      max_water_level: NORMAL_MAX,
      edge_correction: false,
      wlvl_hook: None,
      tie_break: ffi::WS_TIE_FIRST,
      tie_seed: None,
      plots: PlotCfg::none(),
    }
  }

  /// lib.rs:975-985
  #[cfg(feature = "plots")]
  pub const fn set_plot_colour_map(mut self, colour_map: ColourMap) -> Self {
    self.plots.map = Some(colour_map);
    self
  }

  /// lib.rs:991-994: with a folder set, every water level writes `ws_lvl{level}.png` there
  #[cfg(feature = "plots")]
  pub fn set_plot_folder(mut self, path: &std::path::Path) -> Self {
    self.plots.path = Some(path.to_path_buf());
    self
  }

  pub const fn set_max_water_lvl(mut self, max_water_lvl: u8) -> Self {
    self.max_water_level = max_water_lvl;
    self
  }

  pub const fn enable_edge_correction(mut self) -> Self {
    self.edge_correction = true;
    self
  }

  pub const fn set_wlvl_hook(mut self, hook: fn(HookCtx) -> T) -> Self {
    self.wlvl_hook = Some(hook);
    self
  }

  /// Extension: draw the colour of a contested pixel at random among its coloured neighbours like the
  /// reference does (lib.rs:250-253); `seed` makes the draw reproducible.
  pub const fn set_random_tie_break(mut self, seed: Option<u64>) -> Self {
    self.tie_break = ffi::WS_TIE_RANDOM;
    self.tie_seed = seed;
    self
  }

  fn validate(&self, kind: u8) -> Result<(), BuildErr> {
    let cfg = ffi::ws_config {
      kind,
      max_water_level: self.max_water_level,
      edge_correction: self.edge_correction as u8,
      tie_break: self.tie_break,
    };
    match unsafe { ffi::ws_config_validate(&cfg) } {
      ffi::WS_ERR_MAX_TOO_HIGH => Err(BuildErr::MaxToHigh(self.max_water_level)), // lib.rs:1000-1001
      ffi::WS_ERR_MAX_TOO_LOW => Err(BuildErr::MaxToLow(self.max_water_level)),   // lib.rs:1002-1003
      _ => Ok(()),
    }
  }

  pub fn build_merging(self) -> Result<MergingWatershed<T>, BuildErr> {
    self.validate(ffi::WS_MERGING)?;
    Ok(MergingWatershed {
      max_water_level: self.max_water_level,
      edge_correction: self.edge_correction,
      wlvl_hook: self.wlvl_hook,
      tie_break: self.tie_break,
      tie_seed: self.tie_seed,
      plots: self.plots,
    })
  }

  pub fn build_segmenting(self) -> Result<SegmentingWatershed<T>, BuildErr> {
    self.validate(ffi::WS_SEGMENTING)?;
    Ok(SegmentingWatershed {
      max_water_level: self.max_water_level,
      edge_correction: self.edge_correction,
      wlvl_hook: self.wlvl_hook,
      tie_break: self.tie_break,
      tie_seed: self.tie_seed,
      plots: self.plots,
    })
  }
}

/// lib.rs:1051-1065 (the second message names NEVER_FILL, like the reference's)
#[derive(Debug, Clone)]
pub enum BuildErr {
  MaxToHigh(u8),
  MaxToLow(u8),
}

impl std::error::Error for BuildErr {}
impl std::fmt::Display for BuildErr {
  fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
    use BuildErr::*;
    match self {
      MaxToHigh(max) => write!(f, "Maximum water level set to {max}, which is higher than the maximum allowed value {NORMAL_MAX}"),
      MaxToLow(max) => write!(f, "Maximum water level set to {max}, which is lower than the minimum allowed value {NEVER_FILL}"),
    }
  }
}

/// lib.rs:1069-1198
pub trait WatershedUtils {
  /// lib.rs:1081-1087
  fn pre_processor<T, D>(&self, img: nd::ArrayView<T, D>) -> nd::Array<u8, D>
  where
    T: Num + Copy + ToPrimitive + PartialOrd,
    D: nd::Dimension,
  {
    self.pre_processor_with_max::<NORMAL_MAX, T, D>(img)
  }

  /// lib.rs:1134-1173, with every quirk of the reference (min / max folded from zero, `is_normal` gate,
  /// +inf -> ALWAYS_FILL, truncation): the quantisation runs on the device.
  fn pre_processor_with_max<const MAX: u8, T, D>(&self, img: nd::ArrayView<T, D>) -> nd::Array<u8, D>
  where
    T: Num + Copy + ToPrimitive + PartialOrd,
    D: nd::Dimension,
  {
    assert!(MAX < NEVER_FILL); // lib.rs:1143-1144
    assert!(MAX > ALWAYS_FILL);
    let dense = img.as_standard_layout();
    let n = dense.len();
    let mut out = nd::Array::<u8, D>::from_elem(dense.raw_dim(), NEVER_FILL);
    // the element types the engine reads directly; anything else goes through f64, which is what the reference
    // converts every element to anyway (lib.rs:1160) -- the conversion is monotone, so min / max agree
    let dtype = match std::any::type_name::<T>() {
      "f32" => Some(ffi::WS_F32),
      "f64" => Some(ffi::WS_F64),
      "i32" => Some(ffi::WS_I32),
      "u16" => Some(ffi::WS_U16),
      "i16" => Some(ffi::WS_I16),
      "u8" => Some(ffi::WS_U8),
      "i64" => Some(ffi::WS_I64),
      _ => None,
    };
    with_ctx(|ctx| match dtype {
      Some(code) => check(ctx, unsafe {
        ffi::ws_pre_processor(ctx, code, dense.as_ptr() as *const c_void, n, MAX, out.as_mut_ptr())
      }),
      None => {
        let wide: Vec<f64> = dense.iter().map(|x| x.to_f64().unwrap_or(f64::NAN)).collect();
        check(ctx, unsafe {
          ffi::ws_pre_processor(ctx, ffi::WS_F64, wide.as_ptr() as *const c_void, n, MAX, out.as_mut_ptr())
        })
      }
    });
    out
  }

  /// lib.rs:1178-1197: interior pixels strictly GREATER than all 8 neighbours (the code, not its doc comment),
  /// in row-major order.
  fn find_local_minima(&self, img: nd::ArrayView2<u8>) -> Vec<(usize, usize)> {
    with_ctx(|ctx| {
      let mut rc: *mut u64 = std::ptr::null_mut();
      let mut n: usize = 0;
      check(ctx, unsafe { ffi::ws_find_local_minima(ctx, &view(&img), &mut rc, &mut n) });
      let pairs = unsafe { std::slice::from_raw_parts(rc, 2 * n) };
      let out = pairs.chunks_exact(2).map(|p| (p[0] as usize, p[1] as usize)).collect();
      unsafe { ffi::ws_free(rc as *mut c_void) };
      out
    })
  }
}

impl<T> WatershedUtils for MergingWatershed<T> {}
impl<T> WatershedUtils for SegmentingWatershed<T> {}

/// lib.rs:1206-1238
pub trait Watershed<T = ()> {
  fn transform(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> nd::Array2<usize>;
  fn transform_with_hook(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> Vec<T>;
  fn transform_to_list(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> Vec<(u8, Vec<usize>)>;
  fn transform_history(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> Vec<(u8, nd::Array2<usize>)>;
}

/// lib.rs:1297-1305
#[derive(Clone)]
pub struct MergingWatershed<T = ()> {
  max_water_level: u8,
  edge_correction: bool,
  wlvl_hook: Option<fn(HookCtx) -> T>,
  tie_break: u8,
  tie_seed: Option<u64>,
  plots: PlotCfg,
}

/// lib.rs:1609-1617
#[derive(Clone)]
pub struct SegmentingWatershed<T = ()> {
  max_water_level: u8,
  edge_correction: bool,
  wlvl_hook: Option<fn(HookCtx) -> T>,
  tie_break: u8,
  tie_seed: Option<u64>,
  plots: PlotCfg,
}

/// What both transforms share: the configuration that reaches the engine and the four entry points.
struct Engine {
  cfg: ffi::ws_config,
  tie_seed: Option<u64>,
  plots: PlotCfg,
}

impl Engine {
  #[cfg(feature = "plots")]
  fn has_plots(&self) -> bool {
    self.plots.path.is_some()
  }
  #[cfg(not(feature = "plots"))]
  fn has_plots(&self) -> bool {
    false
  }

  /// lib.rs:1472-1487 / 1758-1773: the level's label image without the edge-correction padding; a plot that
  /// fails is reported and the transform goes on.
  #[cfg(feature = "plots")]
  fn draw(&self, ctx: &HookCtx) {
    if let Some(path) = &self.plots.path {
      let shape = ctx.colours.shape();
      let picture = if self.cfg.edge_correction != 0 {
        ctx.colours.slice(nd::s![1..(shape[0] - 1), 1..(shape[1] - 1)])
      } else {
        ctx.colours.view()
      };
      let map: ColourMap = self.plots.map.unwrap_or(plotting::viridis::<usize>); // lib.rs:1011
      if let Err(err) = plotting::plot_slice(picture, &path.join(format!("ws_lvl{}.png", ctx.water_level)), map) {
        println!("Could not make watershed plot. Error: {err}")
      }
    }
  }
  #[cfg(not(feature = "plots"))]
  fn draw(&self, _ctx: &HookCtx) {}

  /// The reference routes `transform`, `transform_history` and `transform_to_list` through
  /// `transform_with_hook` (lib.rs:1524-1560), so they plot too; here they have device paths of their own, and
  /// with a plot folder set one extra pass through the per-level hook makes the pictures.
  fn plot_pass(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) {
    if self.has_plots() {
      self.with_hook(input, seeds, |ctx: HookCtx| self.draw(&ctx));
    }
  }

  fn prepare(&self, ctx: *mut ffi::ws_ctx) {
    if let (ffi::WS_TIE_RANDOM, Some(seed)) = (self.cfg.tie_break, self.tie_seed) {
      check(ctx, unsafe { ffi::ws_ctx_set_tie_seed(ctx, seed) });
    }
  }

  /// Watershed::transform
  fn transform(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> nd::Array2<usize> {
    self.plot_pass(input.view(), seeds);
    // merging: the input shape (lib.rs:1524-1536); segmenting: the padded shape
    let shape = if self.cfg.kind == ffi::WS_MERGING { (input.nrows(), input.ncols()) } else { out_shape(self.cfg.edge_correction != 0, &input) };
    let mut out = nd::Array2::<usize>::zeros(shape);
    let s = seeds_rc(seeds);
    with_ctx(|ctx| {
      self.prepare(ctx);
      check(ctx, unsafe {
        ffi::ws_transform(ctx, &self.cfg, &view(&input), s.as_ptr(), seeds.len(), out.as_mut_ptr() as *mut u64)
      })
    });
    out
  }

  /// Watershed::transform_with_hook with any `FnMut(HookCtx) -> R`: called once per level on this thread, in
  /// level order (lib.rs:1510-1518 / 1796-1804).  A panic inside the hook is carried across the C frames.
  fn with_hook<R>(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)], hook: impl FnMut(HookCtx) -> R) -> Vec<R> {
    struct State<'a, R, F: FnMut(HookCtx) -> R> {
      hook: F,
      out: Vec<R>,
      seeds: &'a [(usize, (usize, usize))],
      panic: Option<Box<dyn std::any::Any + Send>>,
    }
    extern "C" fn tramp<R, F: FnMut(HookCtx) -> R>(user: *mut c_void, h: *const ffi::ws_hook_ctx) {
      let (st, h) = unsafe { (&mut *(user as *mut State<R, F>), &*h) };
      if st.panic.is_some() {
        return;
      }
      let image = unsafe { nd::ArrayView2::from_shape_ptr((h.rows, h.cols), h.image) };
      let colours = unsafe { nd::ArrayView2::from_shape_ptr((h.rows, h.cols), h.colours as *const usize) };
      let ctx = HookCtx { water_level: h.water_level, max_water_level: h.max_water_level, image, colours, seeds: st.seeds };
      let hook = &mut st.hook;
      match std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| hook(ctx))) {
        Ok(v) => st.out.push(v),
        Err(p) => st.panic = Some(p), // never unwind through C
      }
    }
    // lib.rs:1360-1364: colour = index + 1
    let seed_colours: Vec<(usize, (usize, usize))> = seeds.iter().enumerate().map(|(i, &rc)| (i + 1, rc)).collect();
    let mut st = State { hook, out: Vec::new(), seeds: &seed_colours, panic: None };
    let s = seeds_rc(seeds);
    with_ctx(|ctx| {
      self.prepare(ctx);
      check(ctx, unsafe {
        ffi::ws_transform_with_hook(
          ctx,
          &self.cfg,
          &view(&input),
          s.as_ptr(),
          seeds.len(),
          Some(tramp::<R, _>),
          &mut st as *mut _ as *mut c_void,
        )
      })
    });
    if let Some(p) = st.panic.take() {
      std::panic::resume_unwind(p);
    }
    st.out
  }

  /// Watershed::transform_history (lib.rs:1538-1549 / 1824-1835): `(level, colours.to_owned())` per level --
  /// every snapshot is copied once, out of the engine's page-locked buffer into its own Array2.
  fn history(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> Vec<(u8, nd::Array2<usize>)> {
    self.with_hook(input, seeds, |ctx: HookCtx| {
      self.draw(&ctx);
      (ctx.water_level, ctx.colours.to_owned())
    })
  }

  /// Watershed::transform_to_list (lib.rs:1551-1561 / 1837-1847): `(level, find_lake_sizes(colours))` per level.
  /// The rows are `rows * cols + 1` long like the reference's (lib.rs:630); only their first `seeds.len() + 1`
  /// entries can be non-zero, and only those cross the link.
  fn to_list(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> Vec<(u8, Vec<usize>)> {
    extern "C" {
      fn ws_transform_lake_sizes_compact(
        ctx: *mut ffi::ws_ctx,
        cfg: *const ffi::ws_config,
        img: *const ffi::ws_image,
        seeds_rc: *const u64,
        nseeds: usize,
        out_lake_counts: *mut u64,
        out_sizes: *mut u64,
      ) -> c_int;
    }
    self.plot_pass(input.view(), seeds);
    let (r, c) = out_shape(self.cfg.edge_correction != 0, &input);
    let levels = self.cfg.max_water_level as usize + 1;
    let ncol = seeds.len() + 1;
    let mut sizes = vec![0u64; levels * ncol];
    let s = seeds_rc(seeds);
    with_ctx(|ctx| {
      self.prepare(ctx);
      check(ctx, unsafe {
        ws_transform_lake_sizes_compact(ctx, &self.cfg, &view(&input), s.as_ptr(), seeds.len(), std::ptr::null_mut(), sizes.as_mut_ptr())
      })
    });
    let width = ncol.min(r * c + 1);
    (0..levels)
      .map(|l| {
        let mut row = vec![0usize; r * c + 1];
        for (dst, src) in row[..width].iter_mut().zip(&sizes[l * ncol..l * ncol + width]) {
          *dst = *src as usize;
        }
        (l as u8, row)
      })
      .collect()
  }
}

macro_rules! impl_watershed {
  ($name:ident, $kind:expr) => {
    impl<T> $name<T> {
      fn engine(&self) -> Engine {
        Engine {
          cfg: ffi::ws_config {
            kind: $kind,
            max_water_level: self.max_water_level,
            edge_correction: self.edge_correction as u8,
            tie_break: self.tie_break,
          },
          tie_seed: self.tie_seed,
          plots: self.plots.clone(),
        }
      }
    }

    impl<T> Watershed<T> for $name<T> {
      fn transform(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> nd::Array2<usize> {
        self.engine().transform(input, seeds)
      }

      fn transform_with_hook(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> Vec<T> {
        let engine = self.engine();
        match self.wlvl_hook {
          Some(hook) if engine.has_plots() => engine.with_hook(input, seeds, |ctx: HookCtx| {
            engine.draw(&ctx);
            hook(ctx)
          }),
          Some(hook) => engine.with_hook(input, seeds, hook),
          None => {
            // lib.rs:1510, 1520 / 1796, 1806: no hook, no results (without a plot folder the flood is not needed)
            engine.plot_pass(input, seeds);
            Vec::new()
          }
        }
      }

      fn transform_to_list(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> Vec<(u8, Vec<usize>)> {
        self.engine().to_list(input, seeds)
      }

      fn transform_history(&self, input: nd::ArrayView2<u8>, seeds: &[(usize, usize)]) -> Vec<(u8, nd::Array2<usize>)> {
        self.engine().history(input, seeds)
      }
    }
  };
}

impl_watershed!(MergingWatershed, ffi::WS_MERGING);
impl_watershed!(SegmentingWatershed, ffi::WS_SEGMENTING);
