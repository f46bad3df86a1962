//! vectors/src/serializer.rs:1-7
pub trait Serializer {
    fn serialize(&self) -> Vec<u8>;
    fn deserialize(data: Vec<u8>) -> Self;
    fn size(&self) -> usize;
}
