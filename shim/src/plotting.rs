//! The `plots` feature (reference: lib.rs:698-834): one picture of the label image per water level.
//!
//! Visualisation only: the label images come from the engine's per-level hook, the colouring and the PNG
//! encoding run on the host.  The colour-map signature is the reference's
//! (`fn(count, min, max) -> Result<RGBColor, Box<dyn Error>>`, `RGBColor` re-exported from `plotters` so that
//! a user's own colour maps keep their type); the file is written by a small encoder of our own (8-bit RGB,
//! stored deflate blocks), so `plotters` is only needed for that one type.
use ndarray as nd;
use num_traits::ToPrimitive;
use std::{error::Error, io::Write, path::Path};

pub use plotters::style::RGBColor;

#[path = "colour_tables.rs"]
mod colour_tables;

/// pixels with `count <= min` (lib.rs:706, 752-754)
const NAN_COL: RGBColor = RGBColor(0, 0, 0);

/// `((255.0 * count + min) / max) as usize`, clipped to the table (lib.rs:757)
fn index<T: ToPrimitive>(count: &T, min: &T, max: &T) -> Result<usize, Box<dyn Error>> {
  let (c, lo, hi) = (
    count.to_f64().ok_or("colour map: value is not a number")?,
    min.to_f64().ok_or("colour map: min is not a number")?,
    max.to_f64().ok_or("colour map: max is not a number")?,
  );
  let g = (255.0 * c + lo) / hi;
  Ok(if g.is_nan() || g < 0.0 { 0 } else if g > 255.0 { 255 } else { g as usize })
}

/// lib.rs:748-760
pub fn grey_scale<T>(count: T, min: T, max: T) -> Result<RGBColor, Box<dyn Error>>
where
  T: std::fmt::Display + PartialOrd + ToPrimitive,
{
  if count <= min {
    return Ok(NAN_COL);
  }
  let g = index(&count, &min, &max)? as u8;
  Ok(RGBColor(g, g, g))
}

macro_rules! table_map {
  ($name:ident, $table:ident, $cite:literal) => {
    #[doc = $cite]
    pub fn $name<T>(count: T, min: T, max: T) -> Result<RGBColor, Box<dyn Error>>
    where
      T: std::fmt::Display + PartialOrd + ToPrimitive,
    {
      if count <= min {
        return Ok(NAN_COL);
      }
      let [r, g, b] = colour_tables::$table[index(&count, &min, &max)?];
      Ok(RGBColor(r, g, b))
    }
  };
}
table_map!(viridis, VIRIDIS, "lib.rs:762-834 (viridis)");
table_map!(magma, MAGMA, "lib.rs:762-834 (magma)");
table_map!(plasma, PLASMA, "lib.rs:762-834 (plasma)");
table_map!(inferno, INFERNO, "lib.rs:762-834 (inferno)");

fn crc32(bytes: &[u8]) -> u32 {
  let mut c = 0xFFFF_FFFFu32;
  for &b in bytes {
    c ^= b as u32;
    for _ in 0..8 {
      c = if c & 1 != 0 { (c >> 1) ^ 0xEDB8_8320 } else { c >> 1 };
    }
  }
  !c
}

fn adler32(bytes: &[u8]) -> u32 {
  let (mut a, mut b) = (1u32, 0u32);
  for chunk in bytes.chunks(5552) {
    for &x in chunk {
      a += x as u32;
      b += a;
    }
    a %= 65521;
    b %= 65521;
  }
  (b << 16) | a
}

/// 8-bit RGB PNG, `rgb` = `height` rows of `width` pixels (zlib stream of stored blocks)
fn write_png(path: &Path, width: u32, height: u32, rgb: &[u8]) -> Result<(), Box<dyn Error>> {
  let row = 3 * width as usize;
  let mut raw = Vec::with_capacity((row + 1) * height as usize);
  for y in 0..height as usize {
    raw.push(0u8); // filter type 0
    raw.extend_from_slice(&rgb[y * row..(y + 1) * row]);
  }
  let mut z = vec![0x78u8, 0x01];
  if raw.is_empty() {
    z.extend_from_slice(&[1, 0, 0, 0xFF, 0xFF]);
  }
  let nblocks = (raw.len() + 65534) / 65535;
  for (k, block) in raw.chunks(65535).enumerate() {
    let n = block.len() as u16;
    z.push(if k + 1 == nblocks { 1 } else { 0 });
    z.extend_from_slice(&n.to_le_bytes());
    z.extend_from_slice(&(!n).to_le_bytes());
    z.extend_from_slice(block);
  }
  z.extend_from_slice(&adler32(&raw).to_be_bytes());

  fn chunk(file: &mut Vec<u8>, tag: &[u8; 4], data: &[u8]) {
    let mut body = Vec::with_capacity(4 + data.len());
    body.extend_from_slice(tag);
    body.extend_from_slice(data);
    file.extend_from_slice(&(data.len() as u32).to_be_bytes());
    file.extend_from_slice(&body);
    file.extend_from_slice(&crc32(&body).to_be_bytes());
  }
  let mut ihdr = Vec::with_capacity(13);
  ihdr.extend_from_slice(&width.to_be_bytes());
  ihdr.extend_from_slice(&height.to_be_bytes());
  ihdr.extend_from_slice(&[8, 2, 0, 0, 0]); // 8 bits, RGB, deflate, no filter method, no interlace
  let mut file = vec![0x89u8, b'P', b'N', b'G', 0x0D, 0x0A, 0x1A, 0x0A];
  chunk(&mut file, b"IHDR", &ihdr);
  chunk(&mut file, b"IDAT", &z);
  chunk(&mut file, b"IEND", &[]);
  std::fs::File::create(path)?.write_all(&file)?;
  Ok(())
}

/// lib.rs:713-745: min / max folded from the type's default; the picture is `shape[0]` wide and `shape[1]`
/// high with element (x, y) at abscissa x, ordinate y of a cartesian chart (y up).
pub fn plot_slice<'a, T>(
  slice: nd::ArrayView2<'a, T>,
  file_name: &Path,
  color_map: fn(count: T, min: T, max: T) -> Result<RGBColor, Box<dyn Error>>,
) -> Result<(), Box<dyn Error>>
where
  T: Default + std::fmt::Display + std::cmp::PartialOrd + ToPrimitive + Copy,
{
  let mut min = T::default();
  let mut max = T::default();
  for x in slice.iter() {
    if *x < min {
      min = *x;
    }
    if *x > max {
      max = *x;
    }
  }
  let (w, h) = (slice.shape()[0], slice.shape()[1]);
  let mut rgb = vec![0u8; 3 * w * h];
  for ((x, y), px) in slice.indexed_iter() {
    let RGBColor(r, g, b) = color_map(*px, min, max)?;
    let at = 3 * ((h - 1 - y) * w + x);
    rgb[at] = r;
    rgb[at + 1] = g;
    rgb[at + 2] = b;
  }
  write_png(file_name, w as u32, h as u32, &rgb)
}
