//! The reference's README example (README.md:55-67) and a consistency check of the four trait methods, written
//! against the crate's public API only: it compiles against smups/rustronomy-watershed 0.4.1 as well as against
//! this shim.  Needs a B200 and libws_b200.so at run time.
use ndarray as nd;
use ndarray_rand::{rand_distr::Uniform, RandomExt};
use rustronomy_watershed::prelude::*;

#[test]
fn readme_example() {
  let rf = nd::Array2::<u8>::random((512, 512), Uniform::new(0, 254));
  let watershed = TransformBuilder::default().build_segmenting().unwrap();
  let mins = watershed.find_local_minima(rf.view());
  let output = watershed.transform(rf.view(), &mins);
  assert_eq!(output.dim(), (512, 512));
  for (i, &(r, c)) in mins.iter().enumerate() {
    assert_eq!(output[(r, c)], i + 1); // colour = index + 1 (lib.rs:1360-1367)
  }
}

#[test]
fn hook_history_and_list_agree() {
  let rf = nd::Array2::<u8>::random((96, 130), Uniform::new(0, 254));
  let hooked = TransformBuilder::new()
    .set_max_water_lvl(100)
    .set_wlvl_hook(|ctx| (ctx.water_level, ctx.colours.iter().filter(|&&c| c != 0).count()))
    .build_merging()
    .unwrap();
  let mins = hooked.find_local_minima(rf.view());
  let per_level = hooked.transform_with_hook(rf.view(), &mins);
  assert_eq!(per_level.len(), 101);
  let history = hooked.transform_history(rf.view(), &mins);
  let list = hooked.transform_to_list(rf.view(), &mins);
  for (l, ((lvl, coloured), ((hl, img), (ll, sizes)))) in per_level.iter().zip(history.iter().zip(list.iter())).enumerate() {
    assert_eq!((*lvl as usize, *hl as usize, *ll as usize), (l, l, l));
    assert_eq!(img.iter().filter(|&&c| c != 0).count(), *coloured);
    assert_eq!(sizes.len(), 96 * 130 + 1);
    assert_eq!(sizes[1..].iter().sum::<usize>(), *coloured);
  }
  assert!(matches!(TransformBuilder::<()>::new().set_max_water_lvl(255).build_merging(), Err(BuildErr::MaxToHigh(255))));
  assert!(matches!(TransformBuilder::<()>::new().set_max_water_lvl(0).build_segmenting(), Err(BuildErr::MaxToLow(0))));
}
